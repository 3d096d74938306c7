"""CPU: the explained-outlier gate itself (tests/parity_attribution.py) — it must accept a faithful implementation and
refuse an unexplained error.  The 'implementation' here is the oracle's own output, clean or tampered with."""
import pytest
import torch

import parity_attribution as PA
from parity_cases import golden_case

NAME = "e2e_planes_det.npz"


def _oracle_as_implementation(c, tamper=None):
    tr = {}
    out = PA._oracle(c, c["randoms"], tr)
    tr = {k: v.clone() for k, v in tr.items()}
    out = [None if o is None else o.clone() for o in out[:6]]
    if tamper:
        tamper(c, out, tr)
    return out, tr


def _recomposite(c, out, tr, which):
    """recompute the maps of one pass from the (tampered) raw so that the implementation stays self-consistent"""
    _, rd = PA._prep_rays(c)
    cfg = c["opt"].nerf.validation
    z = tr["z_coarse"] if which == "coarse" else tr["z_fine"]
    m = PA._maps_of(tr["raw_" + which], z, rd, cfg, False, None)
    i = 0 if which == "coarse" else 3
    out[i], out[i + 1], out[i + 2] = m["rgb"], m["disp"], m["acc"]
    tr["depth_" + which] = m["depth"]
    if which == "coarse":
        tr["weights_coarse"] = m["weights"]


def test_gate_accepts_the_oracle_itself():
    c = golden_case(NAME)
    out, tr = _oracle_as_implementation(c)
    rep = PA.check_chain(c, "fp32", out, tr, c["randoms"])
    assert rep["coarse_unexplained_max"] == 0.0 and rep["fine_tf_unexplained_max"] == 0.0


def test_gate_refuses_an_unexplained_map_error():
    """a fine-pass sigma error of 0.5e-3 (inside the fp32 raw bound) on every sample is fine; the same maps shifted by
    3e-3 without any cause in the raw values is not"""
    c = golden_case(NAME)

    def shift_maps(c_, out, tr):
        out[3] = out[3] + 3e-3
    out, tr = _oracle_as_implementation(c, shift_maps)
    with pytest.raises(AssertionError, match="fine_tf fp32"):
        PA.check_chain(c, "fp32", out, tr, c["randoms"])


def test_gate_refuses_a_raw_error_above_the_mode_bound():
    c = golden_case(NAME)

    def bump(c_, out, tr):
        tr["raw_coarse"][3, 5, 3] += 0.01        # fp32 sigma bound: 1e-3
        _recomposite(c_, out, tr, "coarse")
    out, tr = _oracle_as_implementation(c, bump)
    with pytest.raises(AssertionError, match="sigma"):
        PA.check_chain(c, "fp32", out, tr, c["randoms"])


def test_gate_attributes_a_last_sample_step_and_nothing_else():
    """flip the sign of a tiny last-sample sigma on a ray with transmittance left: in fp16 mode (sigma bound 0.15) the
    ray's maps jump by T_last and the hybrid must explain it; the same jump WITHOUT the sigma change must fail."""
    c = golden_case(NAME)
    probe = {}

    def step(c_, out, tr):
        raw = tr["raw_fine"]
        w = PA._maps_of(raw, tr["z_fine"], PA._prep_rays(c_)[1], c_["opt"].nerf.validation, False, None)
        cand = torch.nonzero((w["acc"] < 0.7) & (raw[:, -1, 3] <= 0)).flatten()
        r = int(cand[0])
        probe["ray"] = r
        probe["before"] = float(w["acc"][r])
        raw[r, -1, 3] = 0.05                      # the oracle has sigma_last <= 0 there
        _recomposite(c_, out, tr, "fine")
    out, tr = _oracle_as_implementation(c, step)

    # make the oracle's own sigma at that sample small enough to be a legitimate fp16 step: move the ORACLE instead is
    # impossible, so check the two outcomes the gate distinguishes
    tf = {}
    PA._oracle(c, dict(c["randoms"], z_fine=tr["z_fine"]), tf)
    s_ref = float(tf["raw_fine"][probe["ray"], -1, 3])
    if abs(s_ref) <= PA.BOUNDS["fp16"]["sigma"] - 0.05:
        rep = PA.check_chain(c, "fp16", out, tr, c["randoms"])
        assert rep["fine_tf_step_rays"] == 1 and rep["fine_tf_step_raw_max"] > rep["fine_tf_map_bound"]
    else:
        with pytest.raises(AssertionError):       # the sigma change itself exceeds the mode's bound: refused
            PA.check_chain(c, "fp16", out, tr, c["randoms"])


def test_gate_refuses_a_far_index_flip():
    c = golden_case(NAME)

    def flip(c_, out, tr):
        tr["inds"][2, 7] += 3
    out, tr = _oracle_as_implementation(c, flip)
    with pytest.raises(AssertionError, match="index mismatch"):
        PA.check_chain(c, "fp32", out, tr, c["randoms"])
