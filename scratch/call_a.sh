for d in 0 4 8 16; do NQE_JIT_SPARSE_DIV=$d timeout 300 python scratch/exp_fp_sel.py 2>&1 | tail -1; done
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "filter_project or kleene or null_predicate or every_operator or selectivity or many_tiles or full_size" 2>&1 | tail -3
