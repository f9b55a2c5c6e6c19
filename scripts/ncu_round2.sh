#!/bin/bash
# ncu evidence for round 2 (run under gpurun, one GPU): the launch list of the bench command and of one frame pair, and full-set
# captures of the kernels the round changed.  Outputs under gpurun_out/ (summarised into profiles/ by scripts/ncu_summary.py).
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
# (1) launch list of the bench command itself (graph kernel nodes are listed individually; bounded by -c)
CVB_BENCH_ALLPAIRS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_bench_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_launches.log 2>&1
export CVB_NO_GRAPH=1      # kernels appear individually (the extraction otherwise replays one CUDA graph)
# (2) launch list of one frame pair, single context
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches.csv python scripts/prof_kernels.py 1 > gpurun_out/r02_launches.log 2>&1
# (3) full-set captures
for k in ${NCU_KERNELS:-k_ars_sprt k_ars_estimate k_ars_score k_ars_book k_hamming_umma k_fed3 k_suppress_smem k_blur_scharr_pm k_deriv2_v3}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 2 -f -o gpurun_out/r02_$k python scripts/prof_kernels.py 1 > gpurun_out/r02_$k.log 2>&1
  ncu -i gpurun_out/r02_$k.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_pick.py > gpurun_out/r02_$k.txt
done
ls -la gpurun_out | tail -30
