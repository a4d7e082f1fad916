#!/usr/bin/env python
"""Step time of one rank's shard of cfg4 (1/8 of the pods, every node and group) and of the whole cfg4 round with the
lean / wide build of the queue sort forced (BS_SORT_VARIANT is read at bs_create): python profiles/tools/sort_variants.py"""
import importlib, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
pkg = importlib.import_module("batch-scheduler_b200")
S = pkg.snapshot
full = S.config(4)
full.resolve_groups()
shard = full.shard_groups(0, 8)
out = {}
for name, snap in (("shard_1_of_8", shard), ("cfg4", full)):
    for variant in (1, 2, 0):
        os.environ["BS_SORT_VARIANT"] = str(variant)
        eng = pkg.Engine(snap.lanes, 0, fit_bitmap=True, score=True)
        eng.upload(snap)
        for _ in range(5): eng.evaluate_async()
        eng.sync()
        ext = torch.cuda.ExternalStream(eng.stream())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for _ in range(100): eng.evaluate_async()
        e1.record(ext); eng.sync(); torch.cuda.synchronize()
        eng.set_profiling(True); eng.evaluate_async(); eng.sync(); km = eng.kernel_ms(); eng.set_profiling(False)
        out[f"{name}:{('auto', 'lean', 'wide')[variant]}"] = {"step_ms": e0.elapsed_time(e1) / 100, "sort_ms": km["sort"][0], "gang_fit_ms": km["gang_fit"][0]}
        eng.close()
print(json.dumps(out), flush=True)
