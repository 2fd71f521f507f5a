set -x
cd $GRAFT_REPO_ROOT
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
mkdir -p gpurun_out/ncu_counters
M=dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum
for spec in "bc7_amd 8192" "bc7_rg 8192" "bc1 1024" "bc4 4096" "bc5 4096" "bc6h 4096"; do
  set -- $spec
  ncu --metrics $M --clock-control none --csv --log-file gpurun_out/ncu_counters/ctr_$1.csv python tools/ncu_capture.py $1 $2 > gpurun_out/ncu_counters/ctr_$1.log 2>&1
done
python tools/ncu_counters.py bc7_amd=gpurun_out/ncu_counters/ctr_bc7_amd.csv:4194304 bc7_rg=gpurun_out/ncu_counters/ctr_bc7_rg.csv:4194304 bc1=gpurun_out/ncu_counters/ctr_bc1.csv:65536 bc4=gpurun_out/ncu_counters/ctr_bc4.csv:1048576 bc5=gpurun_out/ncu_counters/ctr_bc5.csv:1048576 bc6h=gpurun_out/ncu_counters/ctr_bc6h.csv:1048576 | cut -c1-120
cp profiles/ncu_counters.json gpurun_out/ncu_counters.json
python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err
tail -c 600 gpurun_out/bench_r2_final.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_r2_final_reference.json 2> gpurun_out/bench_r2_final_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_default_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-configs > gpurun_out/ncu_bench.log 2>&1
python -c "
import json
d=json.load(open('gpurun_out/bench_r2_final.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['cpu_baseline']['value'], d['cpu_baseline']['threads_used'], d['parity'])
print({k:d['roofline'][k] for k in ('kernel','frac','traffic','lanes_per_inst','share_of_step')}, d['roofline']['hbm'])
for c in d.get('configs',[]): print(c.get('codec'), round(c.get('value',0),1), round(c.get('e2e',{}).get('value',0),1), c.get('roofline',{}).get('frac'), c.get('clocks',{}).get('samples'), c.get('parity',{}).get('identical_fraction'), c.get('cpu_baseline',{}).get('value'), c.get('error'))
print(json.load(open('gpurun_out/bench_r2_final_reference.json'))['value'])
"
