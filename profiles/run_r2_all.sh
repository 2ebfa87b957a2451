#!/bin/bash
# Round-2 measurement pass on one B200 (gpurun): GPU tests, bench lines of all workloads, launch list, full capture of the
# dominant kernel, FP64 non-FMA peak.  Everything lands in gpurun_out/ and is copied to profiles/r2_* by hand.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_polyhedral.json 2> gpurun_out/bench_polyhedral.err; echo "bench rc=$?"; cat gpurun_out/bench_polyhedral.json
python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cat gpurun_out/bench_reference.json
python profiles/fp64_peak.py > gpurun_out/fp64_peak.log 2>&1; tail -5 gpurun_out/fp64_peak.log
for w in voronoi_4x4x4 dynvor_llzo spherical; do
  python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"; cat gpurun_out/bench_$w.json
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_polyhedral.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-api > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_polyhedral_assign_stream -c 2 -o gpurun_out/stream5 -f \
  python profiles/prof_stream.py 125000 2 > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/stream5.ncu-rep --page raw --csv > gpurun_out/stream5_raw.csv 2>/dev/null
ncu -i gpurun_out/stream5.ncu-rep --page source --csv > gpurun_out/stream5_source.csv 2>/dev/null
ls -la gpurun_out
