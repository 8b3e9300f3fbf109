for d in 0 2 4 8 6 14; do echo "DBG=$d"; GH_GEMM_DBG=$d python tools/gemm_prof.py 2>&1 | sed -n 2,4p | cut -c60-260; done
