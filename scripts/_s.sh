for s in 8 4 8; do
timeout 300 python bench.py --steps 1000 --streams $s --configs none --no-parity > gpurun_out/r2w_s$s.json 2> gpurun_out/r2w_s$s.err
python - <<P
import json; d=json.load(open('gpurun_out/r2w_s$s.json')); e=d['e2e']; print($s, 'value', round(d['value']/1e6,2), round(d['ms_per_step'],5), 'e2e', round(e['value']/1e6,2), 'leader', round(d['leader_chain']['value']/1e6,2), 'wgs84', round(d['wgs84_frame']['value']/1e6,2))
P
done
