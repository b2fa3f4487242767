timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 200 python scripts/dump_root_leafcounts.py gpurun_out/pendulum_root_leafcounts.json | cut -c1-200
cp gpurun_out/pendulum_root_leafcounts.json tests/golden/pendulum_root_leafcounts.json
timeout 400 python bench.py > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err; tail -c 300 gpurun_out/bench_r1_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1_n1.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['config']['leaves_per_step'], d['config']['nodes_per_step'], d['oracle_solves_per_s'], d['cpu_baseline'], d['online_evaluation'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_reference.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_r1_reference.json')); print(d['value'], d['ms_per_step'], d['cpu_baseline'])"
