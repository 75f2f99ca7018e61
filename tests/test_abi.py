"""The C-ABI library builds, loads (no GPU needed) and exports every symbol include/ocbnn_b200.h declares."""
import ctypes
import os
import re

from ocbnn_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ocbnn_b200.h")).read()
    return sorted(set(re.findall(r"OCBNN_API[^;(]*?\b(ocbnn_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) >= 18
    L = ctypes.CDLL(engine.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} declared in the header but not exported"


def test_binding_covers_the_header():
    assert sorted(engine.exported_symbols()) == declared_symbols()
    assert engine.lib().ocbnn_abi_version() == engine.OCBNN_ABI_VERSION
    assert engine.launch_count() >= 0


def test_struct_layout_matches_header():
    """ocbnn_desc is passed by pointer: the ctypes mirror must have the C layout (checked against a compiled probe)."""
    import subprocess, tempfile, textwrap
    src = textwrap.dedent('''
        #include <stdio.h>
        #include <stddef.h>
        #include "ocbnn_b200.h"
        int main(void) {
            printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(ocbnn_desc), offsetof(ocbnn_desc, lik_const), offsetof(ocbnn_desc, prior_mean_vec),
                   offsetof(ocbnn_desc, n_train), offsetof(ocbnn_desc, cparam_stride), offsetof(ocbnn_desc, neg_tau0), offsetof(ocbnn_desc, pos_inv_var));
            return 0;
        }''')
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "p.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "p.c"), "-o", os.path.join(d, "p")])
        got = [int(v) for v in subprocess.check_output([os.path.join(d, "p")]).split()]
    D = engine.ocbnn_desc
    want = [ctypes.sizeof(D), D.lik_const.offset, D.prior_mean_vec.offset, D.n_train.offset, D.cparam_stride.offset, D.neg_tau0.offset,
            D.pos_inv_var.offset]
    assert got == want
