"""Loads tests/golden/round_*.npz fixtures back into Snapshot tables + expected outputs."""
import glob
import importlib
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def fixtures():
    return sorted(glob.glob(os.path.join(HERE, "golden", "round_*.npz")))


def load(path):
    S = importlib.import_module("batch-scheduler_b200.snapshot")
    z = np.load(path)
    def table(cls, prefix):
        return cls(**{f: z[f"{prefix}__{f}"] for f in cls.__dataclass_fields__ if f"{prefix}__{f}" in z.files})
    snap = S.Snapshot(table(S.NodeTable, "nodes"), table(S.PodTable, "pods"), table(S.GroupTable, "groups"),
                      os.path.basename(path))
    out = {k[5:]: z[k] for k in z.files if k.startswith("out__")}
    return snap, out


def replay_fixtures():
    return sorted(glob.glob(os.path.join(HERE, "golden", "replay_*.npz")))


def load_replay(path):
    """(snapshot, queue, expected prefilter / node / ready) of a tests/golden/replay_*.npz fixture."""
    snap, out = load(path)
    return snap, np.load(path)["queue"], out
