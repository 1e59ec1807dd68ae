#!/bin/bash
# tools/profile_round.sh PREFIX -- the ncu evidence of a round, run under gpurun on ONE B200 (profiles/README.md):
#   1. launch list of one cessna 256/16 model (gpu__time_duration per kernel)              -> gpurun_out/PREFIX_launches_cessna256_16.csv
#   2. ncu --set full (+ executed FP32 op counters) of the Level-2 kernels, whole call      -> PREFIX_l2_cessna256_16.ncu-rep, _summary.txt, gpurun_out/traffic.json
#   3. the same for rank 0's share of a 2 / 4 / 8-rank gathering call (GPV_DEBUG_OWN=N,0)   -> PREFIX_l2_share{N}_summary.txt, traffic.json "share"
#   4. launch list of the 10 M-triangle CAD body 1024/2 (where the HBM-bound phases run at size)
# Numbers printed by a run under ncu are never bench values.
set -u
P=${1:-r02}
OUT=gpurun_out/$P
mkdir -p gpurun_out
FP32=smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__thread_inst_executed.sum
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file ${OUT}_launches_cessna256_16.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > ${OUT}_launches.log 2>&1
ncu --set full --metrics $FP32 --clock-control none --import-source on -k regex:"k_l2|k_col_cells" -c 4 -f -o ${OUT}_l2_cessna256_16 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > ${OUT}_full.log 2>&1
python tools/ncu_summary.py ${OUT}_l2_cessna256_16.ncu-rep ${OUT}_l2_cessna256_16 --traffic > /dev/null
for n in 2 4 8; do
  GPV_DEBUG_OWN=$n,0 ncu --set full --metrics $FP32 --clock-control none -k regex:"k_l2|k_col_cells" -c 4 -f -o ${OUT}_l2_share$n python bench.py --steps 1 --warmup 0 --no-cpu-baseline > ${OUT}_share$n.log 2>&1
  python tools/ncu_summary.py ${OUT}_l2_share$n.ncu-rep ${OUT}_l2_share$n --traffic --share $n > /dev/null
done
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file ${OUT}_launches_cad1024_2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --mesh cad --l1 1024 --l2 2 > ${OUT}_launches_cad.log 2>&1
# 5. like-for-like at the reference's own operator boundary: its three kernels (recompiled for sm_100a, strict IEEE) next to the compat tier's, per launch
ncu --target-processes all --metrics gpu__time_duration.sum --clock-control none -k regex:"CUDAClassify|k_compat" --csv --log-file ${OUT}_launches_compat_vs_reference_kernels.csv python tools/bench_reference_gpu_path.py --l1 256 --l2 16 --reps 1 > ${OUT}_compat.log 2>&1
rm -f gpurun_out/*_share*.ncu-rep   # (the whole-call report comes back for the source-level views; the share reports are summarised only)
ls -la gpurun_out/ | tail -20
# back in the build container: cp gpurun_out/traffic.json gpurun_out/${P}_*_summary.txt gpurun_out/${P}_launches_*.csv profiles/
