for d in 78 79; do SN_WG_DBG=$d timeout 300 python scratch/r2/im2col_check.py wgrad_small 2>&1 | grep -v "^  \|Search\|CUDA kernel\|For debugging\|Compile with" | head -6; done
