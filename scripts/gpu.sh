#!/bin/bash
# One parametrised entry point for everything that is run on the GPU box through gpurun:
#   gpurun --timeout 900 -- 'bash scripts/gpu.sh tests bench stages'
# Tasks (any number, run in order; outputs under gpurun_out/, the tail of each is echoed):
#   tests[:expr]     pytest -m gpu (optionally -k expr)          smoke          __graft_entry__.smoke()
#   bench            bench.py default (configs[1])                bench1024 / benchvideo   the other workloads
#   ref              bench.py --impl reference                    stages         per-stage times of one step
#   shapes           per-launch tensor-core table of one UNet forward (scripts/gemm_shapes.py)
#   ab:"o=v,o=v;..." scripts/unet_ab.py A/B of engine options     micro          scripts/micro_bench.py
#   scale:N[:wl]     bench.py on N GPUs (torchrun), workload 512 / 1024 / video
#   ncu:W:K          ncu --set full of the kernels matching K in scripts/ncu_ops.py WHAT=W    trace[:pp]  attention clock64 trace
#   unetlaunches     ncu launch list of one UNet forward
#   launches         ncu launch list (gpu__time_duration) of a bench window -> gpurun_out/launches.txt
#   env: TAG names the output files (default r02), RFB_* are forwarded to the engine as options
mkdir -p gpurun_out
TAG=${TAG:-r02}
for task in "$@"; do
  name=${task%%:*}; arg=""; [[ "$task" == *:* ]] && arg=${task#*:}
  echo "=== $task"
  case $name in
    tests) timeout 1700 python -m pytest tests -m gpu -q -x --timeout 900 --timeout-method=thread ${arg:+-k "$arg"} -s > gpurun_out/${TAG}_pytest.log 2>&1
           echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log; grep -E "rel err|drift|decoded pixel|full path|vae (512|1024)|cond \(|passed|failed|Error|error|rc=" gpurun_out/${TAG}_pytest.log | tail -40 ;;
    smoke) timeout 600 python -u __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log ;;
    bench) timeout 900 python bench.py --steps ${STEPS:-5} --warmup 3 $arg > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 3000 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err ;;
    benchab) # benchab:<OPTION>:<v1>,<v2>[,..]  bench.py with engine option OPTION at each value, interleaved twice (same box)
         opt=${arg%%:*}; IFS=',' read -ra VALS <<< "${arg#*:}"; envname=RFB_$(echo $opt | tr a-z A-Z)
         for r in 1 2; do for v in "${VALS[@]}"; do
           env $envname=$v timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_${opt}${v}_$r.json 2> gpurun_out/${TAG}_bench_${opt}${v}_$r.err
           python -c "import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[2], round(d['value'],3), d['unit'], round(d['ms_per_step'],1), 'ms/step', d['clocks'])" gpurun_out/${TAG}_bench_${opt}${v}_$r.json "$opt=$v" | tee -a gpurun_out/${TAG}_benchab_$opt.txt
         done; done ;;
    bench1024) timeout 1200 python bench.py --workload 1024 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_1024.json 2> gpurun_out/${TAG}_bench_1024.err; tail -c 2500 gpurun_out/${TAG}_bench_1024.json; tail -3 gpurun_out/${TAG}_bench_1024.err ;;
    benchvideo) timeout 1200 python bench.py --workload video --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_video.json 2> gpurun_out/${TAG}_bench_video.err; tail -c 2500 gpurun_out/${TAG}_bench_video.json; tail -3 gpurun_out/${TAG}_bench_video.err ;;
    ref) timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; tail -c 1500 gpurun_out/${TAG}_bench_ref.json ;;
    stages) timeout 600 python scripts/stage_times.py > gpurun_out/${TAG}_stages.log 2>&1; tail -1 gpurun_out/${TAG}_stages.log ;;
    shapes) timeout 600 python scripts/gemm_shapes.py > gpurun_out/${TAG}_shapes${arg:+_$arg}.log 2>&1; head -45 gpurun_out/${TAG}_shapes${arg:+_$arg}.log; tail -1 gpurun_out/${TAG}_shapes${arg:+_$arg}.log ;;
    ab) IFS=';' read -ra SPECS <<< "$arg"; timeout 900 python scripts/unet_ab.py "${SPECS[@]}" > gpurun_out/${TAG}_ab.log 2>&1; tail -20 gpurun_out/${TAG}_ab.log ;;
    unetlaunches) REPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_unet_launches.csv python scripts/unet_once.py > gpurun_out/${TAG}_ncu_unet.log 2>&1
              n=$(grep -o "launches per forward [0-9]*" gpurun_out/${TAG}_ncu_unet.log | grep -o "[0-9]*$"); python scripts/agg_launches.py gpurun_out/${TAG}_unet_launches.csv ${n:-0} > gpurun_out/${TAG}_unet_launches${arg:+_$arg}.txt; head -40 gpurun_out/${TAG}_unet_launches${arg:+_$arg}.txt; rm -f gpurun_out/${TAG}_unet_launches.csv ;;
    ncu) # ncu:<WHAT>:<kernel regex>  one --set full capture of the matching kernels of scripts/ncu_ops.py
         what=${arg%%:*}; kre=${arg#*:}
         WHAT=$what timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kre -c ${COUNT:-2} -o gpurun_out/${TAG}_$what -f python scripts/ncu_ops.py > gpurun_out/${TAG}_ncu_$what.log 2>&1
         python scripts/ncu_summary.py gpurun_out/${TAG}_$what.ncu-rep --traffic gpurun_out/${TAG}_${what}_traffic.json > gpurun_out/${TAG}_${what}_ncu_full.txt 2>&1; cat gpurun_out/${TAG}_${what}_ncu_full.txt | head -60 ;;
    trace) PINGPONG=${arg:-1} timeout 300 python scripts/attn_trace.py > gpurun_out/${TAG}_attn_trace_pp${arg:-1}.txt 2>&1; cat gpurun_out/${TAG}_attn_trace_pp${arg:-1}.txt | cut -c1-260 | head -26 ;;
    scale) # scale:<N>[:<workload>]  bench.py on N GPUs of this box (one rank per GPU, torchrun)
         n=${arg%%:*}; wl=512; [[ "$arg" == *:* ]] && wl=${arg#*:}
         timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --workload $wl --steps ${STEPS:-5} --warmup 3 > gpurun_out/${TAG}_bench_${wl}_${n}gpu.json 2> gpurun_out/${TAG}_bench_${wl}_${n}gpu.err
         tail -c 1800 gpurun_out/${TAG}_bench_${wl}_${n}gpu.json; tail -2 gpurun_out/${TAG}_bench_${wl}_${n}gpu.err ;;
    gemmtrace) timeout 600 python scripts/gemm_trace.py > gpurun_out/${TAG}_gemm_trace.txt 2>&1; cat gpurun_out/${TAG}_gemm_trace.txt | cut -c1-420 ;;
    micro) timeout 900 python scripts/micro_bench.py > gpurun_out/${TAG}_micro.log 2>&1; tail -40 gpurun_out/${TAG}_micro.log ;;
    launches) timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip ${SKIP:-30000} -c ${COUNT:-9000} --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
              python scripts/agg_launches.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches.txt; head -34 gpurun_out/${TAG}_launches.txt; rm -f gpurun_out/${TAG}_launches.csv ;;
    *) echo "unknown task $name" ;;
  esac
done
