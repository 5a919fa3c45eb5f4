#!/bin/bash
# Last GPU call of the round: the K2 variants under lib/var/ (bit-exact tests + short C2 bench each), then the
# fastest passing one through the GPU suite (without the 4-minute statistical parity file: the variants are
# bit-identical, so its outcome cannot change), the full bench line, the per-reach breakdown and the ncu captures.
mkdir -p gpurun_out
V=$PWD/sprfittingpaper2023.jl_b200/lib/var
R=gpurun_out/r2_k2_last_variants.txt
: > $R
for v in x6_signed x5_unsigned x8_signed_nounroll; do
  SPRB200_LIB=$V/$v.so timeout 60 python -m pytest tests/test_gpu_mirror.py tests/test_gpu_api.py -m gpu -x -q > gpurun_out/r2_k2_$v.test.log 2>&1
  rc=$?
  line=$(SPRB200_LIB=$V/$v.so timeout 60 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary 2>/dev/null | tail -1)
  echo "$v $rc $line" | python -c "
import sys, json
v, rc, line = sys.stdin.read().split(' ', 2)
try:
    d = json.loads(line); k = d['roofline']['kernel_ms_per_step']
    print(v, 'tests_rc', rc, 'ms_per_step %.2f' % d['ms_per_step'], 'K2 %.2f' % k['spr_table_kernel'], 'K1 %.2f' % k['spr_traj_kernel'])
except Exception as e:
    print(v, 'tests_rc', rc, 'ms_per_step 1e9 bench failed', e)
" >> $R
  tail -1 $R
done
BEST=$(python -c "
rows=[l.split() for l in open('$R') if l.strip()]
ok=[(float(r[4]), r[0]) for r in rows if r[2]=='0' and float(r[4])<1e8]
print(min(ok)[1] if ok else 'x5_unsigned')
")
echo "chosen $BEST" | tee -a $R
export SPRB200_LIB=$V/$BEST.so
timeout 60 python -m pytest tests -m gpu -x -q --ignore=tests/test_gpu_parity.py > gpurun_out/r2_pytest_last_noparity.log 2>&1; echo "suite (without the parity file) rc=$?"; tail -2 gpurun_out/r2_pytest_last_noparity.log
timeout 150 python bench.py > gpurun_out/r2_bench_last.json 2> gpurun_out/r2_bench_last.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2_bench_last.json
timeout 30 python scripts/c2_breakdown.py > gpurun_out/r2_c2_breakdown_last.txt 2>&1; tail -12 gpurun_out/r2_c2_breakdown_last.txt
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_last.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/r2_ncu_launch_last.log 2>&1; echo "ncu launches rc=$?"
timeout 60 ncu --set full --clock-control none --import-source on -k regex:spr_table_kernel -c 1 -f -o gpurun_out/r2_k2_last_reach35 python scripts/ncu_c2_class.py 9 20 > gpurun_out/r2_ncu_k2_last.log 2>&1; echo "ncu K2 rc=$?"
