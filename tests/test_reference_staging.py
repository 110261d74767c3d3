"""The staged reference (baseline/_ref, used by `bench.py --impl reference` and the GPU
reference bar) is the UNMODIFIED reference, and the loop body driven on it (oracle/ref_loop.py)
is the one the golden vectors were recorded from and the oracle restates."""
import json
import os

import pytest
import torch

import common
import helpers
from oracle import progan_oracle as O
from oracle import ref_loop, stage_reference


def test_staged_copies_are_byte_identical_to_the_reference():
    d = stage_reference.staged_dir()
    if d is None and os.path.isdir(stage_reference.REF):
        d = stage_reference.stage()
    if d is None:
        pytest.skip("no staged reference and no /root/reference on this machine")
    man = json.load(open(os.path.join(d, "MANIFEST.json")))["sha256"]
    for name in stage_reference.FILES:
        assert stage_reference.sha256(os.path.join(d, name)) == man[name]
        live = os.path.join(stage_reference.REF, name)
        if os.path.exists(live):
            assert stage_reference.sha256(live) == man[name], name


def test_loop_on_real_reference_equals_the_oracle_iteration():
    """One full iteration (both Adam steps and the EMA) of the loop body on the real reference
    modules against oracle.train_iteration on the same seeded weights and inputs."""
    R, where = ref_loop.import_reference()
    if R is None:
        pytest.skip("reference modules not available")
    inp = common.make_inputs("s2_a0.5")
    G, D, Grun, g_opt, d_opt = ref_loop.build(R, inp["channel"], inp["z_dim"])
    G.load_state_dict(inp["G"]); D.load_state_dict(inp["D"]); Grun.load_state_dict(inp["G"])
    dl, gp, gl = ref_loop.iteration(G, D, Grun, g_opt, d_opt, inp["real"], inp["z"], inp["eps"], 2, 0.5)
    PG, PD, PR = O.params_of(inp["G"]), O.params_of(inp["D"]), O.params_of(inp["G"], False)
    res = O.train_iteration(PG, PD, PR, O.AdamState(PG), O.AdamState(PD), inp["real"], inp["z"],
                            inp["eps"], 2, 0.5)
    assert helpers.rel(gp, res["grad_penalty"]) < 2e-5
    assert helpers.rel(dl, res["disc_loss"]) < 2e-5 and helpers.rel(gl, res["gen_loss"]) < 2e-5
    for k, p in D.named_parameters():
        assert helpers.rel(p, PD[k]) < 1e-5, k
    for k, p in Grun.named_parameters():
        assert helpers.rel(p, PR[k]) < 1e-5, k
