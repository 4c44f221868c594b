"""GPU: the seeded random scenes of tests/test_fuzz_oracle_vs_ref.py (there: reference == oracle, bit for bit) through
the CUDA path against the oracle.  Depth, coverage and alpha identical; colour within 1 LSB per channel (north_star's
stated tolerance); DoF-R pixel-identical given identical inputs, so with DoF the same 1-LSB bound is checked on the
blurred frame against the oracle's DoF of ITS frame only where the GPU's unblurred frame equals the oracle's."""
import numpy as np
import pytest

from swegl_b200 import _abi, configs

pytestmark = pytest.mark.gpu


def channel_diff(a, b):
    a8 = a.view(np.uint8).reshape(a.shape + (4,)).astype(np.int16)
    b8 = b.view(np.uint8).reshape(b.shape + (4,)).astype(np.int16)
    return np.abs(a8 - b8)


def gpu_frame(renderer, scene, vp, screen):
    renderer.upload_scene(scene)
    renderer.set_screen(*screen)
    renderer.begin_frame(scene)
    px = np.zeros((screen[1], screen[0]), np.uint32)
    z = np.empty((vp.h, vp.w), np.float32)
    st = renderer.render(vp, px, z)
    return px, z, st


@pytest.mark.parametrize("seed", list(range(48)) + [60])    # 60: a lower half ending at (int)ceil(+huge) == INT_MIN (k_setup plan_slot)
def test_fuzz_scene_matches_oracle(renderer, oracle, seed):
    scene, vp, screen, pose = configs.fuzz_case(seed)
    post = vp.post_mode
    vp.post_mode = _abi.POST_NULL
    px, z, st = gpu_frame(renderer, scene, vp, screen)
    o = oracle.render(scene, vp, screen_wh=screen)
    assert (z.view(np.uint32) == o["z"].view(np.uint32)).all(), "depth buffer differs"
    d = channel_diff(px, o["pixels"])
    assert d[..., 3].max() == 0 and d.max() <= 1
    assert int(st.n_covered) == o["n_covered"]
    if post == _abi.POST_DOF:
        vp.post_mode = post
        bpx, bz, _ = gpu_frame(renderer, scene, vp, screen)
        bo = oracle.render(scene, vp, screen_wh=screen)
        assert (bz.view(np.uint32) == bo["z"].view(np.uint32)).all()
        if (px == o["pixels"]).all():
            assert (bpx == bo["pixels"]).all()             # integer box average of identical inputs
        else:
            assert channel_diff(bpx, bo["pixels"]).max() <= 1   # an average of values each within 1 is within 1


@pytest.mark.parametrize("seed", list(range(12)))
def test_fuzz_scene_with_transparency_layers_matches_oracle(renderer, oracle, seed):
    scene, vp, screen, pose = configs.fuzz_layers_case(seed)
    px, z, st = gpu_frame(renderer, scene, vp, screen)
    o = oracle.render(scene, vp, screen_wh=screen)
    assert (z.view(np.uint32) == o["z"].view(np.uint32)).all()
    d = channel_diff(px, o["pixels"])
    # a blended pixel combines up to 1 + L shaded colours, each within 1 LSB, with weights summing to <= 1
    assert d[..., 3].max() == 0 and d.max() <= 1


@pytest.mark.parametrize("seed,n_bands", [(3, 2), (7, 3), (9, 5)])
def test_fuzz_scene_bands_reassemble(renderer, seed, n_bands):
    """sort-first row bands of a fuzz frame, culling on, equal the unsplit frame bit for bit"""
    scene, vp, screen, pose = configs.fuzz_case(seed)
    full, fz, _ = gpu_frame(renderer, scene, vp, screen)
    px = np.zeros_like(full)
    z = np.empty_like(fz)
    cuts = [round(k * vp.h / n_bands) for k in range(n_bands + 1)]
    try:
        for b0, b1 in zip(cuts, cuts[1:]):
            vp.band = (b0, b1)
            renderer.render(vp, px, z)
    finally:
        vp.band = (0, 0)
    assert (px == full).all() and (z.view(np.uint32) == fz.view(np.uint32)).all()
