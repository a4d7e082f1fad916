"""CPU: the C-ABI library loads and exports every symbol include/bsched.h declares; flag and enum
values agree between the header, the ctypes binding, the snapshot module and the oracle header.
No compute call is made (there is no GPU here and no CPU path in the product)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header():
    return open(os.path.join(ROOT, "include", "bsched.h")).read()


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.capi.load()
    hdr = _header()
    declared = set(re.findall(r"\b(bs_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"bs_engine"}
    assert declared, "no declarations parsed"
    assert declared == set(pkg.capi.SYMBOLS), (declared ^ set(pkg.capi.SYMBOLS))
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.bs_abi_version() == int(re.search(r"#define BS_ABI_VERSION (\d+)", hdr).group(1))


def test_no_device_fails_loudly(pkg):
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    with pytest.raises(pkg.capi.BsError) as ei:
        pkg.Engine(5)
    assert ei.value.code == pkg.capi.BS_E_NODEVICE


def test_flag_values_agree(pkg, snapshot_mod):
    hdr = _header()
    ohdr = open(os.path.join(ROOT, "oracle", "bs_oracle.h")).read()

    def defs(text, prefix):
        return {m.group(1): int(m.group(2), 0) for m in
                re.finditer(r"#define %s([A-Z_]+)\s+\(?(-?0x[0-9a-fA-F]+|-?\d+)u?\)?" % prefix, text)}
    h, o = defs(hdr, "BS_"), defs(ohdr, "BSO_")
    for k in ("NODE_NIL", "NODE_NO_NODE", "NODE_UNSCHEDULABLE", "NODE_TAINTS_ERR", "POD_PERMITTED_RECENTLY",
              "POD_OCC_NOREFS", "POD_OCC_MISMATCH", "POD_LISTER_MISS", "GROUP_SCHEDULED", "GROUP_HAS_POD",
              "GROUP_HAS_MINRES", "GROUP_DENIED"):
        assert h[k] == o[k] == getattr(snapshot_mod, k), k
    assert h["MAX_LANES"] == o["MAX_LANES"] == snapshot_mod.MAX_LANES
    assert snapshot_mod.GID_NONE == -1 and snapshot_mod.GID_MISSING == -2
    assert "#define BS_GID_NONE (-1)" in hdr and "#define BS_GID_MISSING (-2)" in hdr
    # enums
    for name, val in (("BS_PF_ERR_NOT_ENOUGH", snapshot_mod.PF_NOT_ENOUGH), ("BS_UNSCHEDULABLE", snapshot_mod.UNSCHEDULABLE),
                      ("BS_CODE_WAIT", pkg.capi.CODE_WAIT), ("BS_E_REF_PANIC", pkg.capi.BS_E_REF_PANIC)):
        m = re.search(r"%s = (-?\d+)" % name, hdr)
        assert m and int(m.group(1)) == val, name


def test_product_never_references_oracle():
    # the product path must not import / link / call anything under oracle/
    pkgdir = os.path.join(ROOT, "batch-scheduler_b200")
    for dp, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "bs_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f
