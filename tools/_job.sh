mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_page_parts.py -m gpu -q -k "many_column or item_driven" 2>&1 | tail -2
timeout 800 bash tools/sanitize.sh memcheck "parts or malformed or merge or overlapping or bool or crc_on_read or verify_on_read or item_driven or many_column or small or tombstone or predicates or pruning"
timeout 600 bash tools/sanitize.sh racecheck "item_driven or many_column or c4_shape_cut or merge_tables or verify_on_read or scan_bool"
