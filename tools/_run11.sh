set -x
timeout 600 python -m pytest tests/test_deconv_gpu.py tests/test_prologue_gpu.py -x -q 2>&1 | tail -3
timeout 300 python tools/decoder_bench.py > gpurun_out/r2z_decoder_bench.json 2>gpurun_out/r2z_decoder_bench.err; head -5 gpurun_out/r2z_decoder_bench.json; tail -3 gpurun_out/r2z_decoder_bench.err
timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_ov.json
python - <<'PY'
import json
s=open('gpurun_out/bench_ov.json').read(); j=json.loads(s[s.index('{'):])
print(j['ms_per_step'], j['e2e']['ms_per_step'])
PY
