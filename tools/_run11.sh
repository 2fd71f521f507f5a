set -x
cd $GRAFT_REPO_ROOT
timeout 2400 python -m pytest tests -q -m gpu --durations=6 2>&1 | tail -30
mkdir -p gpurun_out/ncu_counters
M=dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum
for spec in "bc7_amd 8192 4194304" "bc7_rg 8192 4194304" "bc1 1024 65536" "bc4 4096 1048576" "bc5 4096 1048576" "bc6h 4096 1048576"; do
  set -- $spec
  ncu --metrics $M --clock-control none --csv --log-file gpurun_out/ncu_counters/ctr_$1.csv python tools/ncu_capture.py $1 $2 > gpurun_out/ncu_counters/ctr_$1.log 2>&1
done
python tools/ncu_counters.py bc7_amd=gpurun_out/ncu_counters/ctr_bc7_amd.csv:4194304 bc7_rg=gpurun_out/ncu_counters/ctr_bc7_rg.csv:4194304 bc1=gpurun_out/ncu_counters/ctr_bc1.csv:65536 bc4=gpurun_out/ncu_counters/ctr_bc4.csv:1048576 bc5=gpurun_out/ncu_counters/ctr_bc5.csv:1048576 bc6h=gpurun_out/ncu_counters/ctr_bc6h.csv:1048576 | cut -c1-200
cp profiles/ncu_counters.json gpurun_out/ncu_counters.json
timeout 900 python bench.py --steps 3 > gpurun_out/bench_r2_default.json 2> gpurun_out/bench_r2_default.err
tail -c 1500 gpurun_out/bench_r2_default.err
python -c "
import json
d=json.load(open('gpurun_out/bench_r2_default.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
print(d['roofline'])
for c in d.get('configs',[]): print(c.get('codec'), c.get('value'), c.get('e2e',{}).get('value'), c.get('roofline',{}).get('frac'), c.get('roofline',{}).get('traffic'), c.get('clocks',{}).get('samples'), c.get('error'))
"
