#!/bin/bash
# One parameterised GPU visit (replaces the per-experiment scripts of round 1).  Usage, under gpurun:
#
#   scripts/gpu_visit.sh <tag> <step> [<step> ...]
#
# Steps (each writes gpurun_out/<tag>_<step>.log and prints its last lines):
#   tests[:<pytest -k expr>]   python -m pytest tests -m gpu
#   smoke                      __graft_entry__.smoke()
#   bench[:ENV=V,ENV=V]        python bench.py --steps 10 --warmup 3 --no-cpu-baseline with the environment given (A/B switches)
#   benchargs:--a,v,--b,w      the same with extra bench.py arguments (commas for spaces), e.g. benchargs:--mesh,wild
#   fullbench                  the default bench line incl. the CPU baseline leg, and the reference arm
#   train | synth | metrics | birnn   the other bench workloads
#   launches[:workload]        ncu launch list (gpu__time_duration) of two bench steps -> <tag>_launches[_workload].csv
#   prof:<name>:<kernel regex>:<skip>[:workload]   ncu --set full of ONE launch -> <tag>_prof_<name>.ncu-rep
#   profall                    after `launches`: one ncu --set full capture of every kernel kind of the step (scripts/find_skips.py)
#   micro:<MxNxK>[,...]        scripts/gemm_microbench.py
#   sass                       opcode histogram of the tcgen05 executor and the fan kernel from the shipped .so
tag=${1:-visit}; shift
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for step in "$@"; do
  name=${step%%:*}; arg=""; [ "$step" != "$name" ] && arg=${step#*:}
  log=gpurun_out/${tag}_${name}.log
  case $name in
    tests)
      rm -f gpurun_out/parity_report.jsonl
      if [ -n "$arg" ]; then timeout -s KILL 1500 python -m pytest tests -q -m gpu -k "$arg" > $log 2>&1; else timeout -s KILL 1500 python -m pytest tests -q -m gpu > $log 2>&1; fi
      echo "rc=$?" >> $log; tail -n 12 $log
      [ -f gpurun_out/parity_report.jsonl ] && cp gpurun_out/parity_report.jsonl gpurun_out/${tag}_parity_report.jsonl ;;
    smoke)
      timeout -s KILL 300 python __graft_entry__.py --smoke > $log 2>&1; echo "rc=$?" >> $log; tail -n 3 $log ;;
    bench)
      suffix=$(echo "$arg" | tr -c 'A-Za-z0-9=\n' '_'); log=gpurun_out/${tag}_bench_${suffix}.log
      env $(echo "$arg" | tr ',' ' ') timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-companions > $log 2>&1; echo "rc=$?" >> $log
      echo "bench [$arg]"; tail -n 2 $log | cut -c1-700 ;;
    benchargs)
      suffix=$(echo "$arg" | tr -c 'A-Za-z0-9=\n' '_'); log=gpurun_out/${tag}_bench_${suffix}.log
      timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-companions $(echo "$arg" | tr ',' ' ') > $log 2>&1; echo "rc=$?" >> $log
      echo "bench args [$arg]"; tail -n 2 $log | cut -c1-400 ;;
    fullbench)
      timeout -s KILL 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.log 2>&1
      timeout -s KILL 900 python bench.py --steps 20 --warmup 3 > $log 2>&1; echo "rc=$?" >> $log; tail -n 2 $log | cut -c1-1500 ;;
    train|synth|metrics|birnn)
      timeout -s KILL 600 python bench.py --workload $name --steps 10 --warmup 3 $arg > $log 2>&1; echo "rc=$?" >> $log; tail -n 2 $log | cut -c1-600 ;;
    launches)
      wl=""; suffix=""; [ -n "$arg" ] && wl="--workload $arg" && suffix="_$arg"
      timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${tag}_launches${suffix}.csv \
          python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-companions $wl > $log 2>&1; echo "rc=$?" >> $log; tail -n 1 $log | cut -c1-200 ;;
    prof)
      pname=${arg%%:*}; rest=${arg#*:}; regex=${rest%%:*}; rest=${rest#*:}; skip=${rest%%:*}; wl=""; [ "$rest" != "$skip" ] && wl="--workload ${rest#*:}"
      timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -o gpurun_out/${tag}_prof_$pname -f \
          python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-companions $wl > gpurun_out/${tag}_prof_$pname.log 2>&1; echo "prof $pname rc=$?" ;;
    profall)
      # one `ncu --set full` capture per kernel kind of the step, located with the launch list of THIS visit (run `launches` first)
      python scripts/find_skips.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_skips.txt
      while read pname regex skip; do
        timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -o gpurun_out/${tag}_prof_$pname -f \
            python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-companions > gpurun_out/${tag}_prof_$pname.log 2>&1; echo "prof $pname (skip $skip) rc=$?"
      done < gpurun_out/${tag}_skips.txt ;;
    micro)
      timeout -s KILL 300 python scripts/gemm_microbench.py $(echo "$arg" | tr ',' ' ') > $log 2>&1; grep -h tflops $log | tr -d '\n'; echo ;;
    sass)
      cuobjdump -sass em-pose_b200/libempose_b200.so 2>/dev/null | python scripts/sass_histogram.py > $log; tail -n 40 $log ;;
    *) echo "unknown step $step" ;;
  esac
done
ls gpurun_out | grep "^${tag}_" | head -60
