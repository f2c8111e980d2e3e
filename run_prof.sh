set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "mapping_rows or qual_seq or best_hits or resident" 2>&1 | tail -15 > gpurun_out/r02w_tests.log
