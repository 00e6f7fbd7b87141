#!/bin/bash
# ncu --set full captures of the hot-path kernels (one gpurun call).  Usage: TAG=r2k bash profiles/ncu_all.sh
cd ${GRAFT_REPO_ROOT:-.}
TAG=${TAG:-r2}
export STEP_CONF=${STEP_CONF:-4096}
python profiles/ncu_targets.py > gpurun_out/${TAG}_targets_plain.log 2>&1 || { cat gpurun_out/${TAG}_targets_plain.log; exit 1; }
cat gpurun_out/${TAG}_targets_plain.log | grep -v Warning
cap() {  # name regex skip target
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o gpurun_out/${TAG}_$1 python profiles/ncu_targets.py $4 > gpurun_out/${TAG}_$1.log 2>&1
  tail -1 gpurun_out/${TAG}_$1.log
}
if [ -z "$ONLY_FULL" ]; then
STEP_PASSES=1 cap filter resident_filter_kernel 0 c3   # first filter launch: all structures, full degree
STEP_PASSES=1 cap lanczos resident_lanczos_kernel 0 c3
STEP_PASSES=1 cap build16 resident_build16_kernel 0 c3
STEP_PASSES=1 cap spmm64 spmm_paired_kernel 0 c3
STEP_PASSES=1 cap gram2 gram2_dmma_kernel 0 c3
STEP_PASSES=1 cap rotate rotate_resid_dmma_kernel 0 c3
STEP_PASSES=1 cap rr rr_kernel 1 c3
STEP_PASSES=1 cap contacts contacts_tiled_kernel 1 c3
STEP_PASSES=1 cap assemble assemble_rows_kernel 0 c3
cap dcc gemm_nt_dmma_kernel 1 dcc
cap slab dense_slab_apply_kernel 1 slab
cap tf32 dense_slab_tf32_kernel 1 tf32
fi
cap sytrd sytrd_kernel 1 full          # the whole tridiagonalisation is ONE launch
cap dcgemm dgemm_dmma_kernel 63 full  # 57 GEMM launches per call (7 merge levels + Gram + T_b V_b^T + 24 x 2): #63 = top-level merge of the second call
ls -la gpurun_out/${TAG}_*.ncu-rep
