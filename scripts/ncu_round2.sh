#!/bin/bash
# ncu evidence for round 2 (run under gpurun, one GPU): the launch list of the bench command and of one frame pair, and full-set
# captures of the kernels the round changed.  Only the exported summaries are written under gpurun_out/ (the .ncu-rep files stay in
# /tmp on the box: gpurun copies back at most 64 MiB).
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out /tmp/ncu
T=${NCU_TAG:-r02}
if [ "${NCU_SKIP_LISTS:-0}" != "1" ]; then
# (1) launch list of the bench command itself (graph kernel nodes are listed individually; bounded by -c)
CVB_BENCH_ALLPAIRS=0 CVB_BENCH_TRACK=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_bench_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_bench_launches.log 2>&1
# (2) launch list of one frame pair, single context, kernels launched one by one
CVB_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches.csv python scripts/prof_kernels.py 1 > gpurun_out/${T}_launches.log 2>&1
fi
export CVB_NO_GRAPH=1
# (3) full-set captures: name[:skip[:count]]
for spec in ${NCU_KERNELS:-k_ars_sprt:1:1 k_ars_estimate8:0:1 k_ars_estimate8:3:1 k_ars_score:0:1 k_ars_score:4:1 k_ars_book:2:1 k_ars_begin:1:1 k_hamming_umma:2:1 k_fed3:0:2 k_suppress_smem:1:1 k_blur_scharr_pm:0:1 k_deriv2_v3:0:1}; do
  IFS=: read k s c <<< "$spec"; s=${s:-2}; c=${c:-1}
  o=${T}_${k}_s${s}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -f -o /tmp/ncu/$o python scripts/prof_kernels.py 1 > gpurun_out/$o.log 2>&1
  ncu -i /tmp/ncu/$o.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_pick.py > gpurun_out/$o.txt
  ncu -i /tmp/ncu/$o.ncu-rep --page details 2>/dev/null | grep -v "^ *$" | head -400 > gpurun_out/$o.details.txt
  ncu -i /tmp/ncu/$o.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/$o.source.csv.gz
  tail -2 gpurun_out/$o.log > gpurun_out/$o.log.tail; mv gpurun_out/$o.log.tail gpurun_out/$o.log
done
du -sh gpurun_out; ls -la gpurun_out | tail -50
