for rep in 1 2; do
for cfg in "" "DAI_DEC_CHUNK=9600" "DAI_DEC_CHUNK=19200" "DAI_TC_TWO_PASS=25" "DAI_DEC_CHUNK=9600 DAI_TC_TWO_PASS=25"; do
  tag=$(echo "$cfg" | tr ' =' '__'); [ -z "$tag" ] && tag=default
  env $cfg timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r02l_${tag}_$rep.json 2>/dev/null
done
done
env DAI_DEC_CHUNK=19200 timeout 300 python bench.py --no-extras --no-cpu-baseline --roots 32 > gpurun_out/r02l_R32_chunk19200.json 2>/dev/null
timeout 300 python bench.py --no-extras --no-cpu-baseline --roots 32 > gpurun_out/r02l_R32_default.json 2>/dev/null
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02l_*.json')):
    try:
        d = json.load(open(f)); print(f, round(d['value'], 1), {k: round(v, 2) for k, v in d['roofline'].get('step_share_ms', {}).items()})
    except Exception as e: print(f, 'ERR', e)
PY
