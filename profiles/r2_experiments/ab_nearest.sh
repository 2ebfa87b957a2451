for lib in site_analysis_b200/libsab200.so build/variants/lib_nb3.so build/variants/lib_nb4.so; do
  for w in voronoi_4x4x4 dynvor_llzo; do
    SAB200_LIB=$PWD/$lib python bench.py --workload $w --no-cpu-baseline --no-api --steps 3 > gpurun_out/ab_n.json 2>/dev/null
    python -c "import json; d=json.load(open('gpurun_out/ab_n.json')); print('$lib', '$w', d['ms_per_step'], d['roofline']['kernel_ms'], d['parity']['mismatches'])"
  done
done
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k 'spherical or speculative or large_overlapping' 2>&1 | tail -2
SAB_DEBUG=1 python bench.py --workload spherical --no-cpu-baseline --no-api --steps 3 > gpurun_out/sph.json 2> gpurun_out/sph.err; grep resolve gpurun_out/sph.err | tail -1
python -c "import json; d=json.load(open('gpurun_out/sph.json')); print('spherical', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['parity'])"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/poly20.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/poly20.json')); print('polyhedral', d['value'], d['ms_per_step'], d['value_with_history'], d['e2e_api'])"
