export PYTHONPATH=$PWD
for c in 2,16,16,64,64,3,1,1,1,0 2,16,16,32,32,3,1,1,1,0 4,160,160,32,32,3,1,1,1,0 2,40,40,128,128,3,1,1,0,0 3,24,24,128,64,3,1,1,0,0 2,80,80,128,64,3,1,1,0,0 2,48,48,64,64,3,1,0,0,1; do
  timeout 60 python tools/gpu_conv_selftest.py --one $c 2>&1 | grep -E "^\["
done
python tools/gpu_layer_times.py 64 640 > gpurun_out/layers4.log 2>&1; head -2 gpurun_out/layers4.log
python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['stage_ms'], d['roofline']['achieved'], d['roofline']['frac'])"
