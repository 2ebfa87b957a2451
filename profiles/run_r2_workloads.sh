python profiles/fp64_peak.py > gpurun_out/fp64_peak.log 2>&1; cat gpurun_out/fp64_peak.log
for w in voronoi_4x4x4 dynvor_llzo spherical; do python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -2 gpurun_out/bench_$w.err; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_$w.json')); print('$w', {k:d[k] for k in ('value','ms_per_step','e2e','e2e_api','value_with_history','roofline','cpu_baseline','parity')})"; done
