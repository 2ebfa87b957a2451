#!/bin/bash
# Round-2 final measurement pass on one B200 (gpurun): all GPU tests, bench lines of every workload + the reference arm,
# launch list of the headline command, full captures of the dominant kernel of every workload.  Output: gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_polyhedral.json 2> gpurun_out/bench_polyhedral.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
for w in voronoi_4x4x4 dynvor_llzo spherical; do
  python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"
done
SAB_RESOLVE_FAST=1 python bench.py --workload spherical --no-cpu-baseline --no-api > gpurun_out/bench_spherical_per_frame_resolver.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 200 --csv --log-file gpurun_out/launches_polyhedral.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-api --parity-frames 0 > gpurun_out/ncu_launch.log 2>&1
for spec in "voronoi_4x4x4:k_nearest_assign_cells" "dynvor_llzo:k_dynvor_assign_cells" "spherical:k_spherical_assign_cells" "spherical:k_resolve_spec"; do
  w=${spec%%:*}; k=${spec#*:}
  ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/full_$k -f \
    python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline --no-api --parity-frames 0 > gpurun_out/ncu_$k.log 2>&1
  ncu -i gpurun_out/full_$k.ncu-rep --page raw --csv > gpurun_out/full_${k}_raw.csv 2>/dev/null
  rm -f gpurun_out/full_$k.ncu-rep
done
ls -la gpurun_out | tail -30
