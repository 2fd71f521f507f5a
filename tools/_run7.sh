set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_bc7amd.py tests/test_golden.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python tools/amd_mode_times.py 2048 2>&1 | tail -20
cat > /tmp/opq.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
import gfx_imagecompress_b200 as g
from gfx_imagecompress_b200 import synth
g.load_library(); g.init(0)
n = int(sys.argv[1]); kind = sys.argv[2]; mask = int(sys.argv[3], 0)
dev = torch.device("cuda", 0)
px = torch.from_numpy(synth.rgba8_gradnoise(n, n, 3, kind)).to(dev)
out = torch.empty((n * n // 16, 16), dtype=torch.uint8, device=dev)
o = g.Opts.default(amd_mode_mask=mask)
for _ in range(2):
    g.encode_device(g.BC7_AMD, px, synth.FMT_RGBA8, n, n, 1, opts=o, out=out)
torch.cuda.synchronize()
PY
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_opaque1024.csv python /tmp/opq.py 1024 opaque 0xff > gpurun_out/ncu_a.log 2>&1
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_ramp1024.csv python /tmp/opq.py 1024 ramp 0xff > gpurun_out/ncu_b.log 2>&1
