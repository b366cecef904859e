# The single-GPU measurements behind profiles/bench_r2_*.json, r2_launches.csv and r2_ncu_group_kernel.txt (run under gpurun from the repo root).
set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/f_n1.json 2> gpurun_out/f_n1.err
python bench.py --steps 20 --warmup 3 --init box --no-cpu-baseline > gpurun_out/f_n1_box.json 2>/dev/null
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_ref.json 2>/dev/null
for c in c1 c3 c5; do python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/f_$c.json 2>/dev/null; done
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/f_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/f_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lnprob_group -s 8 -c 2 -o gpurun_out/f_prof_group -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/f_ncu2.log 2>&1
tail -2 gpurun_out/f_ncu2.log
python -c "
import json
for f in ('f_n1','f_n1_box','f_ref','f_c1','f_c3','f_c5'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d.get('value'), d.get('e2e',{}).get('value'), d.get('ms_per_step'))
    except Exception as e: print(f, 'ERR', e)
"
