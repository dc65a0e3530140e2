"""The CUDA path against the reference Engine's own listings (see tests/test_qa_decks.py), and its print-cycle balances against
the oracle's to rounding."""
import numpy as np
import pytest
import torch
import qa_decks
from test_qa_decks import check_elem_samp, check_ct3a, check_inibri

pytestmark = pytest.mark.gpu
if torch.cuda.is_available():
    from openradioss_b200.engine import Engine
    from oracle.orc import Oracle


def test_elem_samp_gpu_reproduces_the_reference_listing():
    rg = check_elem_samp(Engine(qa_decks.elem_samp()))
    ro = check_elem_samp(Oracle(qa_decks.elem_samp()))
    assert np.abs(rg[:, 1:6] - ro[:, 1:6]).max() <= 1e-9 * np.abs(ro[:, 1:6]).max()      # 2045 cycles of both


def test_ct3a_gpu_reproduces_the_reference_listing():
    check_ct3a(Engine(qa_decks.ct3a()))


def test_inibri_stress_gpu_reproduces_the_reference_listing():
    check_inibri(Engine(qa_decks.inibri_stress()))


def test_history_ring_matches_the_per_cycle_queries():
    g = Engine(qa_decks.elem_samp()); g.set_print(True)
    rows = []
    for _ in range(5):
        g.run_cycles(1); b = g.balance(); rows.append([b[k] for k in ("encin", "enrot", "enint", "wfext", "xmomt", "ymomt", "zmomt", "xmass")])
    g.run_cycles(20)
    h = g.balance_history(25)
    assert np.array_equal(h[:5], np.array(rows))
    assert h[-1, 2] > h[4, 2] > 0.0
