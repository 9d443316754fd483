#!/bin/bash
# One GPU session: tests, benches, then ncu captures (each ncu only after the same command exited 0 without it).
# usage: tools/gpu_round.sh [tests] [bench] [ncu_layers] [ncu_k1]
mkdir -p gpurun_out
for what in "$@"; do
case $what in
tests)
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest_exit=$?"; tail -15 gpurun_out/pytest_gpu.log;;
bench)
  timeout 600 python bench.py --tiles ${TILES:-262144} --steps 5 --warmup 3 > gpurun_out/bench_tiles.json 2> gpurun_out/bench_tiles.err; echo "bench_exit=$?"; tail -2 gpurun_out/bench_tiles.err; cat gpurun_out/bench_tiles.json
  timeout 600 python bench.py --workload varint1g --steps 5 --warmup 3 > gpurun_out/bench_varint1g.json 2> gpurun_out/bench_varint1g.err; echo "bench2_exit=$?"; tail -2 gpurun_out/bench_varint1g.err; cat gpurun_out/bench_varint1g.json;;
benchfull)
  timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "benchfull_exit=$?"; tail -2 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json;;
fixtures)
  timeout 600 python bench.py --workload fixtures --steps 5 --warmup 3 > gpurun_out/bench_fixtures.json 2> gpurun_out/bench_fixtures.err; echo "benchfx_exit=$?"; tail -2 gpurun_out/bench_fixtures.err; cat gpurun_out/bench_fixtures.json;;
ncu_layers)
  CMD="python bench.py --tiles 65536 --steps 1 --warmup 3 --no-cpu-baseline"
  $CMD > gpurun_out/plain_layers.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k "regex:k_decode_class|k_assemble_layers|k0_|k_finalize|scan_" -s 54 -c 18 -f -o gpurun_out/prof_layers $CMD > gpurun_out/ncu_layers.log 2>&1
  echo "ncu_layers_exit=$?"; tail -3 gpurun_out/ncu_layers.log;;
ncu_k1)
  CMD="python bench.py --workload varint1g --stream-bytes 268435456 --steps 1 --warmup 3 --no-cpu-baseline"
  $CMD > gpurun_out/plain_k1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k "regex:k1a_|k1b_" -s 4 -c 2 -f -o gpurun_out/prof_k1 $CMD > gpurun_out/ncu_k1.log 2>&1
  echo "ncu_k1_exit=$?"; tail -3 gpurun_out/ncu_k1.log;;
traffic)
  # dram bytes per launch of the dominant kernels at the bench's own sizes (profiles/traffic.json)
  CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
  timeout 600 $CMD > gpurun_out/plain_traffic_tiles.log 2>&1 && \
  timeout 1500 ncu --set full --clock-control none -k "regex:k_assemble_layers" -s 3 -c 1 -f -o gpurun_out/prof_traffic_tiles $CMD > gpurun_out/ncu_traffic_tiles.log 2>&1
  echo "traffic_tiles_exit=$?"; tail -2 gpurun_out/ncu_traffic_tiles.log
  CMD="python bench.py --workload varint1g --steps 1 --warmup 3 --no-cpu-baseline"
  timeout 600 $CMD > gpurun_out/plain_traffic_k1.log 2>&1 && \
  timeout 1500 ncu --set full --clock-control none -k "regex:k1a_|k1b_" -s 6 -c 2 -f -o gpurun_out/prof_traffic_k1 $CMD > gpurun_out/ncu_traffic_k1.log 2>&1
  echo "traffic_k1_exit=$?"; tail -2 gpurun_out/ncu_traffic_k1.log;;
exp)
  # sweep environment variables (comma-separated names share the value): EXP_VAR=COVT_ASM_MINB EXP_VALUES="1 10 12" tools/gpu_round.sh exp
  for val in $EXP_VALUES; do
    echo "== $EXP_VAR=$val"
    ENVS=""; for name in ${EXP_VAR//,/ }; do ENVS="$ENVS $name=$val"; done
    env $ENVS timeout 600 python bench.py --tiles ${TILES:-262144} --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/exp_tiles_$val.err | python tools/kernel_times.py
    env $ENVS timeout 600 python bench.py --workload fixtures --replicas 64 --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/exp_fx_$val.err | python tools/kernel_times.py
  done;;
launches)
  CMD="python bench.py --tiles 65536 --steps 1 --warmup 3 --no-cpu-baseline"
  $CMD > gpurun_out/plain_launches.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
  echo "launches_exit=$?";;
esac
done
