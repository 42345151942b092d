for r in 1 2 3; do for d in 0 7; do CRW_SEG_R=$r CRW_SEG_DBG=$d timeout 120 python tools/dbg_sp_T.py 2>&1 | grep "T=8" | sed "s/^/R=$r dbg=$d /"; done; done
for r in 2 4 8; do for d in 7; do CRW_SEG_NCH128=1 CRW_SEG_R=$r CRW_SEG_DBG=$d timeout 120 python tools/dbg_sp_T.py 2>&1 | grep "T=8" | sed "s/^/nch128 R=$r dbg=$d /"; done; done
