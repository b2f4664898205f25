bash tools/sanitize.sh memcheck
bash tools/sanitize.sh racecheck
bash tools/capture_profiles.sh r02
