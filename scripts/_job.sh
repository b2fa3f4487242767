timeout 600 python -m pytest tests/test_gpu_explicit_mpc.py tests/test_gpu_batch_parity.py -x -q -m gpu 2>&1 | tail -6
timeout 400 python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; tail -c 400 gpurun_out/bench_r1_final.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1_final.json')); print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['frac'], d['roofline']['achieved'], d['cpu_baseline'])"
timeout 200 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_r1_final_reference.json 2>/dev/null; cut -c1-300 gpurun_out/bench_r1_final_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1_final_launches_root22.csv python scripts/run_partition.py --roots 22 > /dev/null 2>&1; wc -l gpurun_out/r1_final_launches_root22.csv
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_lcss_eval -s 9 -c 1 -o gpurun_out/prof_eval_r1_final python scripts/run_partition.py --roots 0,1,2,3,4,5,6,7,8,9,10,11 > gpurun_out/prof_eval_final.log 2>&1; tail -1 gpurun_out/prof_eval_final.log | cut -c1-100
