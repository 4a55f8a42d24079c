#!/usr/bin/env python
"""Block-level timeline of consecutive control cycles (config 2, 65,536 Pandas): when do the blocks of cycle c + 1 start relative
to the end of cycle c, and how full are the SM slots over time?  Uses osc_debug_block_times (globaltimer stamps at the start and
end of every block of the fused kernel).  Run on the GPU box:  python tools/pipeline_trace.py >> profiles/r02_pipeline_trace.md
(SAI_B200_NO_PIPELINE=1 for the grid-wide wait)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import sai_primitives_b200 as sp  # noqa: E402
from sai_primitives_b200 import batched, capi  # noqa: E402
import bench  # noqa: E402

batched.JointTask.DefaultParameters.use_internal_otg = False
batched.MotionForceTask.DefaultParameters.use_internal_otg = False


def run(label):
    lib = capi.load_library()
    R = 65536
    q, dq, _, rng = bench.sample_batch(sp, R, 0)
    robot = sp.BatchedRobot("panda", R)
    robot.setQ(q); robot.setDq(dq); robot.updateModel()
    mft = sp.MotionForceTask(robot, bench.LINK, (np.eye(3), np.array(bench.POINT))); jt = sp.JointTask(robot)
    ctrl = sp.RobotController(robot, [mft, jt])
    dev = torch.device("cuda", 0)
    d_q = torch.from_numpy(np.ascontiguousarray(q.T)).to(dev); d_dq = torch.from_numpy(np.ascontiguousarray(dq.T)).to(dev)
    d_tau = torch.zeros((7, R), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    for _ in range(40):
        ctrl.stepDevice(d_q.data_ptr(), d_dq.data_ptr(), d_tau.data_ptr())
    robot.sync()
    blocks = lib.osc_debug_block_times(robot.handle, 1, None, 0)
    for _ in range(24):          # the last 8 of these are kept
        ctrl.stepDevice(d_q.data_ptr(), d_dq.data_ptr(), d_tau.data_ptr())
    robot.sync()
    buf = np.zeros(blocks * 16, dtype=np.uint64)
    assert lib.osc_debug_block_times(robot.handle, 1, buf.ctypes.data_as(C.c_void_p), buf.size) == blocks
    t = buf.reshape(8, blocks, 2).astype(np.int64)
    order = np.argsort(t[:, :, 0].min(axis=1))
    t = t[order]
    t0 = t.min()
    print("### %s\n" % label)
    print("| cycle | first block starts | last block starts | first block ends | last block ends | blocks of the NEXT cycle started before this one ended |")
    print("|---:|---:|---:|---:|---:|---:|")
    for c in range(8):
        s, e = t[c, :, 0] - t0, t[c, :, 1] - t0
        nxt = int((t[c + 1, :, 0] - t0 < e.max()).sum()) if c < 7 else -1
        print("| %d | %.1f us | %.1f us | %.1f us | %.1f us | %s |" % (c, s.min() / 1e3, s.max() / 1e3, e.min() / 1e3, e.max() / 1e3, "-" if nxt < 0 else "%d of %d" % (nxt, blocks)))
    # resident blocks over time (slots = 2 per SM), sampled every 100 ns between the start of cycle 1 and the end of cycle 6
    lo, hi = t[1, :, 0].min(), t[6, :, 1].max()
    ts = np.arange(lo, hi, 100)
    starts = np.sort(t[:, :, 0].reshape(-1)); ends = np.sort(t[:, :, 1].reshape(-1))
    resident = np.searchsorted(starts, ts, side="right") - np.searchsorted(ends, ts, side="right")
    slots = 2 * torch.cuda.get_device_properties(0).multi_processor_count
    span = (t[6, :, 0].min() - t[1, :, 0].min()) / 5.0 / 1e3
    print("\nmean resident blocks %.1f of %d slots (%.1f %%), time with fewer than half of the slots busy %.1f %%, cycle period %.2f us (%.3g cycles/s)\n"
          % (resident.mean(), slots, 100.0 * resident.mean() / slots, 100.0 * (resident < slots / 2).mean(), span, R / (span * 1e-6)))
    robot.close()


if __name__ == "__main__":
    mode = "grid-wide wait (SAI_B200_NO_PIPELINE=1)" if os.environ.get("SAI_B200_NO_PIPELINE") == "1" else "cross-cycle pipelining (default)"
    run(mode)
