#!/bin/bash
# One gpurun call: GPU tests, the bench line, engine / NLMPC probes, ncu launch list and full captures.  Outputs under gpurun_out/.
# usage: gpurun --timeout 2400 -- 'bash tools/gpu_round.sh [tests] [bench] [probe] [nlmpc] [launches] [ncu_lmpc] [ncu_nlmpc]'
mkdir -p gpurun_out
want() { [[ " $ARGS " == *" $1 "* ]]; }
ARGS="$*"
[ -z "$ARGS" ] && ARGS="tests bench probe nlmpc launches ncu_lmpc ncu_nlmpc"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt
if want tests; then
  timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
  tail -3 gpurun_out/pytest_gpu.log
fi
if want lmpc_tests; then
  timeout 900 python -m pytest tests/test_gpu_lmpc.py tests/test_gpu_lmpc_properties.py tests/test_gpu_golden.py tests/test_gpu_closed_loop.py -x -q > gpurun_out/pytest_lmpc.log 2>&1; echo "pytest lmpc rc=$?"
  tail -5 gpurun_out/pytest_lmpc.log
fi
if want nlmpc_tests; then
  timeout 1500 python -m pytest tests/test_gpu_nlmpc.py tests/test_gpu_nlmpc_solve.py tests/test_gpu_nlmpc_structured.py tests/test_gpu_nlmpc_closed_loop.py tests/test_gpu_baseline_configs.py tests/test_gpu_user_systems.py tests/test_gpu_golden.py -x -q > gpurun_out/pytest_nlmpc.log 2>&1; echo "pytest nlmpc rc=$?"
  tail -5 gpurun_out/pytest_nlmpc.log
fi
if want quick; then
  timeout 300 python tools/bshort.py default > gpurun_out/bshort.txt 2>&1; cat gpurun_out/bshort.txt
fi
if want bench; then
  timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
  cut -c1-600 gpurun_out/bench.json
  timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
fi
if want probe; then
  timeout 600 python tools/engine_probe.py ${PROBE_BATCHES:-1 148 592 4096} > gpurun_out/engine_probe.jsonl 2>&1; echo "probe rc=$?"
  cut -c1-1500 gpurun_out/engine_probe.jsonl
fi
if want nlmpc; then
  timeout 900 python tools/bench_nlmpc.py > gpurun_out/nlmpc_bench.jsonl 2>&1; echo "nlmpc rc=$?"
  cut -c1-400 gpurun_out/nlmpc_bench.jsonl
fi
if want launches; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1; echo "launches rc=$?"
fi
if want probe2; then
  ( CTA_THREADS=384 timeout 300 python tools/engine_probe.py 148; B200MPC_CTA_PIPE=1 timeout 300 python tools/engine_probe.py 148; \
    B200MPC_CTA_PIPE=1 CTA_THREADS=384 timeout 300 python tools/engine_probe.py 148 ) > gpurun_out/engine_probe2.jsonl 2>&1; echo "probe2 rc=$?"
  cut -c1-1500 gpurun_out/engine_probe2.jsonl
fi
if want sweep; then
  # BASELINE configs[4]: horizon sweep at the stated global batch (here on however many GPUs the call has; 1 GPU = one shard of 65536)
  for PH in 10 20 50 100; do
    timeout 600 python bench.py --global-batch 65536 --ph $PH --steps 3 --warmup 3 --no-nlmpc --no-cpu-baseline >> gpurun_out/horizon_sweep.jsonl 2>> gpurun_out/sweep.err
  done
  echo "sweep rc=$?"; cut -c1-200 gpurun_out/horizon_sweep.jsonl
fi
if want ncu_lmpc; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:lmpc_solve_kernel -s 2 -c 1 -o gpurun_out/lmpc_full -f \
    python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-nlmpc > gpurun_out/ncu_lmpc.log 2>&1; echo "ncu_lmpc rc=$?"
fi
if want ncu_cta; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:lmpc_cta_kernel -s 1 -c 1 -o gpurun_out/lmpc_cta_full -f \
    python tools/cta_one.py ${CTA_BATCH:-148} > gpurun_out/ncu_lmpc_cta.log 2>&1; echo "ncu_lmpc_cta rc=$?"
fi
if want ncu_nlmpc; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:nlmpc_structured_kernel -s 0 -c 1 -o gpurun_out/nlmpc_struct_full -f \
    python tools/nlmpc_one.py ugv30 1024 > gpurun_out/ncu_nlmpc.log 2>&1; echo "ncu_nlmpc rc=$?"
fi
# reports are too large to travel back together (64 MiB limit): export the pages we read as CSV, keep a report only if KEEP_REP names it
for rep in gpurun_out/*.ncu-rep; do
  [ -f "$rep" ] || continue
  b="${rep%.ncu-rep}"
  ncu -i "$rep" --page raw --csv > "${b}_raw.csv" 2>/dev/null
  ncu -i "$rep" --page source --csv --print-source sass > "${b}_sass.csv" 2>/dev/null
  ncu -i "$rep" --page details --csv > "${b}_details.csv" 2>/dev/null
  case " $KEEP_REP " in *" $(basename "$b") "*) ;; *) rm -f "$rep";; esac
done
ls -la gpurun_out
du -sh gpurun_out
