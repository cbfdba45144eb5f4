for WT in 4 5 6; do
MSNAP_WRITE_MINB=$WT python bench.py --steps 100 --configs cfg5 --no-parity --no-cpu-baseline > gpurun_out/r2_b7.json 2>gpurun_out/r2_b7.err; python -c "
import json; d=json.load(open('gpurun_out/r2_b7.json')); print('MINB=$WT', d['value'], d['ms_per_step'])
for k,v in d['variants'].items():
    if 'kernels_ms_per_step' in v: print(k, '%.4g'%v['value'], v['ms_per_step'], {a:round(b,3) for a,b in v['kernels_ms_per_step'].items() if 'write' in a or 'count' in a}, v['rows_per_step'])"
tail -3 gpurun_out/r2_b7.err
done
