"""Experiment: the batch split into G groups of clips, each on its own CUDA stream, so that one group's IK / affinity
kernels fill the SMs another group's ALS launch leaves idle in its tail. Prints frames/s for G = 1, 2, 3."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from multiview_motion_capture_b200.clips import ClipBatch

B, K, W = 1184, 5, 7
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
kps, n_pose, Kc, RT, _ = bench.make_inputs(B, W + K + 1, seed=1000, distinct=148)
kps_dev = torch.from_numpy(kps).to(dev)
np_dev = torch.from_numpy(n_pose).to(dev)
for G in (1, 2, 3, 4):
    per = B // G
    cbs, streams = [], []
    for g in range(G):
        cb = ClipBatch(per, bench.N_VIEWS, bench.N_PEOPLE, max_tracks=40, max_new=bench.N_PEOPLE, device=dev)
        cb.set_calib(Kc[g * per:(g + 1) * per], RT[g * per:(g + 1) * per])
        cbs.append(cb)
        streams.append(torch.cuda.Stream(dev))
    torch.cuda.synchronize()
    def step(f):
        for g in range(G):
            with torch.cuda.stream(streams[g]):
                cbs[g].step_device(kps_dev[f, g * per:(g + 1) * per], np_dev[f, g * per:(g + 1) * per], f)
    for s in range(W):
        step(1 + s)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d = torch.cuda.current_stream(dev)
    ev0.record(d)
    for st in streams:
        st.wait_stream(d)
    for s in range(K):
        step(1 + W + s)
    for st in streams:
        d.wait_stream(st)
    ev1.record(d)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    print(f"groups {G}: {G * per * K / (ms * 1e-3):.1f} frames/s  ({ms / K:.1f} ms per step of {G * per} clips)", flush=True)
    for cb in cbs:
        cb.close()
