"""The gradient-comparison rule of tests/test_gpu_reference_dropin.py (`_check_grads`) on synthetic tensors, no GPU: what it
accepts (differences inside the reference's own run-to-run noise; ill-conditioned tensors with float64 evidence) and what it
must keep rejecting (a well-conditioned tensor that is off, anything beyond 1e-4)."""
import pytest
import torch

import test_gpu_reference_dropin as T


def _runs(want, noise_rel, seed):
    g = torch.Generator().manual_seed(seed)
    return {k: w + noise_rel.get(k, 0.0) * float(w.abs().max()) * (2 * torch.rand(w.shape, generator=g) - 1) for k, w in want.items()}


@pytest.fixture()
def grads():
    g = torch.Generator().manual_seed(0)
    return {"enc.weight": torch.randn(8, 4, 3, 3, generator=g), "enc.bias": 1e-9 * torch.randn(8, generator=g),
            "dec.weight": 0.1 * torch.randn(4, 8, 3, 3, generator=g)}


def test_inside_the_reference_noise_passes(grads):
    noise = {"dec.weight": 3e-4}                                  # the reference differs from itself by 3e-4 here (atomics)
    runs = [_runs(grads, noise, s) for s in (1, 2, 3)]
    got = _runs(grads, noise, 9)                                  # ours: a draw from the same distribution
    T._check_grads(got, grads, runs)
    with pytest.raises(AssertionError):                           # the same difference on a tensor that has NO noise is an error
        T._check_grads(_runs(grads, {"enc.weight": 3e-4}, 9), grads, runs)


def test_bias_noise_is_measured_against_the_layer(grads):
    got = dict(grads, **{"enc.bias": grads["enc.bias"] + 5e-9})   # rounding noise of a bias whose true gradient is zero
    T._check_grads(got, grads, [grads])


def test_float64_evidence_only_helps_ill_conditioned_tensors(grads):
    got = dict(grads, **{"enc.weight": grads["enc.weight"] + 3e-5 * float(grads["enc.weight"].abs().max())})
    exact = {k: w.double() for k, w in grads.items()}
    with pytest.raises(AssertionError):                           # the reference's float32 value is exact: no excuse
        T._check_grads(got, grads, [grads], truth=lambda: exact)
    loose = dict(exact, **{"enc.weight": exact["enc.weight"] + 1e-3 * float(grads["enc.weight"].abs().max())})
    T._check_grads(got, grads, [grads], truth=lambda: loose)      # the reference itself is 1e-3 from its float64 value
    far = dict(grads, **{"enc.weight": grads["enc.weight"] + 2e-4 * float(grads["enc.weight"].abs().max())})
    with pytest.raises(AssertionError):                           # beyond 1e-4 nothing helps
        T._check_grads(far, grads, [grads], truth=lambda: loose)


def test_failed_evidence_run_reports_the_miss(grads):
    got = dict(grads, **{"enc.weight": grads["enc.weight"] + 3e-5 * float(grads["enc.weight"].abs().max())})

    def boom():
        raise RuntimeError("no float64 run")
    with pytest.raises(AssertionError, match="float64 run of the reference failed"):
        T._check_grads(got, grads, [grads], truth=boom)
