for s in 2 3 4; do
timeout 300 python bench.py --steps 300 --streams $s --configs none --no-parity > gpurun_out/r2w_b$s.json 2> gpurun_out/r2w_b$s.err
python - <<P
import json; d=json.load(open('gpurun_out/r2w_b$s.json')); e=d['e2e']; print($s, 'value', d['value'], d['ms_per_step'], 'e2e', e['value'], e['ms_per_step'], e['d2h_GBps_per_gpu'], 'all', e['all_outputs']['value'])
P
done
