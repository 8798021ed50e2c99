mkdir -p gpurun_out
B="--no-e2e --no-c5 --no-fp64 --no-cpu-baseline"
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/r2_env30.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest30.log 2>&1; tail -3 gpurun_out/r2_gputest30.log
timeout 600 python bench.py > gpurun_out/r2_bench30.json 2> gpurun_out/r2_bench30.err; tail -c 300 gpurun_out/r2_bench30.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:narrow_kernel -s 3 -c 1 -o gpurun_out/r2_narrow_c5_v2 python scripts/c5_probe.py 270 75 0 > gpurun_out/r2_ncu30a.log 2>&1; tail -2 gpurun_out/r2_ncu30a.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 3 -c 1 -o gpurun_out/r2_tile_f64_v2 python scripts/kernel_probe.py 24 f64 tile:11 > gpurun_out/r2_ncu30b.log 2>&1; tail -2 gpurun_out/r2_ncu30b.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:lec_ -c 400 --csv --log-file gpurun_out/r2_launches30.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_ncu30c.log 2>&1; tail -c 200 gpurun_out/r2_ncu30c.log
python -c "
import json
d=json.load(open('gpurun_out/r2_bench30.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['clocks'])
print('e2e', d['e2e']['value'], d['e2e']['packed_int16']['value'], d['e2e']['plugin']['value'])
print('fp64', d['fp64']['value'], d['fp64']['roofline']['frac'])
print('c5', d['c5']['steps_per_s'], d['c5']['frac'], d['c5']['roofline']['frac'], d['c5']['roofline']['kernel_ms'], d['c5']['roofline']['finalize_kernel_ms'])
"
