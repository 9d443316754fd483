#!/bin/bash
# One GPU session: tests, benches, then ncu captures (each ncu only after the same command exited 0 without it).
# usage: [R=r02_z] tools/gpu_round.sh [tests] [benchfull] [reference] [fixtures] [varint] [launches] [ncu_layers] [ncu_k1] [traffic]
# Everything lands in gpurun_out/${R}_*; copy what should be judged into profiles/.
R=${R:-r02_z}
mkdir -p gpurun_out
for what in "$@"; do
case $what in
tests)
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest_exit=$?"; tail -4 gpurun_out/${R}_pytest_gpu.log;;
benchfull)
  timeout 900 python bench.py > gpurun_out/${R}_bench_config5_1Mtiles.json 2> gpurun_out/${R}_bench_full.err; echo "benchfull_exit=$?"; tail -2 gpurun_out/${R}_bench_full.err; python tools/kernel_times.py < gpurun_out/${R}_bench_config5_1Mtiles.json;;
reference)
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference_arm_config5.json 2> gpurun_out/${R}_bench_ref.err; echo "reference_exit=$?"; cut -c1-400 gpurun_out/${R}_bench_reference_arm_config5.json;;
fixtures)
  timeout 600 python bench.py --workload fixtures --steps 5 --warmup 3 > gpurun_out/${R}_bench_config2_fixtures_x256.json 2> gpurun_out/${R}_bench_fx.err; echo "fx_exit=$?"; python tools/kernel_times.py < gpurun_out/${R}_bench_config2_fixtures_x256.json
  timeout 600 python bench.py --workload fixtures --rle-topology --steps 5 --warmup 3 > gpurun_out/${R}_bench_config2_fixtures_x256_rle_topology.json 2> gpurun_out/${R}_bench_fxr.err; echo "fxr_exit=$?"; python tools/kernel_times.py < gpurun_out/${R}_bench_config2_fixtures_x256_rle_topology.json
  timeout 600 python bench.py --workload fixtures --props --steps 5 --warmup 3 > gpurun_out/${R}_bench_config2_fixtures_x256_props.json 2> gpurun_out/${R}_bench_fxp.err; echo "fxp_exit=$?"; python tools/kernel_times.py < gpurun_out/${R}_bench_config2_fixtures_x256_props.json;;
varint)
  timeout 600 python bench.py --workload varint1g --steps 5 --warmup 3 > gpurun_out/${R}_bench_config3_varint1g.json 2> gpurun_out/${R}_bench_v1g.err; echo "varint_exit=$?"; python tools/kernel_times.py < gpurun_out/${R}_bench_config3_varint1g.json;;
ncu_layers)
  CMD="python bench.py --tiles 262144 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e-host"
  $CMD > gpurun_out/${R}_plain_layers.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k "regex:k_decode_class|k_assemble_layers|k0_|k_tile_status|k_layer_totals|scan_" -s 57 -c 19 -f -o gpurun_out/${R}_prof_layers $CMD > gpurun_out/${R}_ncu_layers.log 2>&1
  echo "ncu_layers_exit=$?"; tail -3 gpurun_out/${R}_ncu_layers.log
  python tools/ncu_summary.py gpurun_out/${R}_prof_layers.ncu-rep > gpurun_out/${R}_layers_262144tiles_ncu_summary.txt;;
ncu_k1)
  CMD="python bench.py --workload varint1g --steps 1 --warmup 3 --no-cpu-baseline"
  $CMD > gpurun_out/${R}_plain_k1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k "regex:k1a_|k1b_" -s 6 -c 2 -f -o gpurun_out/${R}_prof_k1 $CMD > gpurun_out/${R}_ncu_k1.log 2>&1
  echo "ncu_k1_exit=$?"; tail -3 gpurun_out/${R}_ncu_k1.log
  python tools/ncu_summary.py gpurun_out/${R}_prof_k1.ncu-rep > gpurun_out/${R}_k1_two_pass_1GiB_ncu_summary.txt;;
traffic)
  # dram bytes per launch of the dominant kernel at the bench's own size (profiles/traffic.json)
  CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e-host"
  timeout 600 $CMD > gpurun_out/${R}_plain_traffic_tiles.log 2>&1 && \
  timeout 1500 ncu --set full --clock-control none -k "regex:k_assemble_layers" -s 3 -c 1 -f -o gpurun_out/${R}_prof_traffic_tiles $CMD > gpurun_out/${R}_ncu_traffic_tiles.log 2>&1
  echo "traffic_tiles_exit=$?"; tail -2 gpurun_out/${R}_ncu_traffic_tiles.log
  python tools/ncu_summary.py gpurun_out/${R}_prof_traffic_tiles.ncu-rep > gpurun_out/${R}_assemble_1Mtiles_ncu_summary.txt; cat gpurun_out/${R}_assemble_1Mtiles_ncu_summary.txt;;
launches)
  CMD="python bench.py --no-cpu-baseline --no-e2e-host"
  $CMD > gpurun_out/${R}_plain_launches.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${R}_launches_default_bench_command_first2000.csv $CMD > gpurun_out/${R}_ncu_launches.log 2>&1
  echo "launches_exit=$?";;
esac
done
