"""CPU: the pin of DoF-R.  post_shader_depth_box as shipped cannot produce an image (post_shaders.hpp:63-111 reads a depth
buffer opaque fragments never write and indexes the source out of bounds), so the oracle's DoF-R is
checked against the reference built with oracle/dof_r.patch -- 7 changed lines in that one header, every other source the
reference's own (oracle/Makefile target refdofr).  swegl::render() of the patched build, DoF included, must equal the
oracle's frame bit for bit: the named DoF workloads, and seeded fuzz scenes over focal distances / depths that put the
blur radius through all of 0..5 (and the focal_depth == 1 branch of remap_clipped, lerp.hpp:33)."""
import os

import numpy as np
import pytest

from swegl_b200 import _abi, configs


@pytest.fixture(scope="module")
def ref_dofr():
    from oracle.binding import Ref, REF_DOFR_LIB
    if not os.path.exists(REF_DOFR_LIB):
        pytest.skip("oracle/_ref/libswegl_ref_dofr.so not built (needs /root/reference; run `make -C oracle refdofr`)")
    return Ref(REF_DOFR_LIB)


def ref_frame(ref, scene, vp, screen, pose):
    h = ref.import_scene(scene)
    scr = ref.lib.ref_screen_new(*screen)
    rv = ref.make_viewport(scr, vp, pose, with_dof=True)
    px, z = ref.render(h, rv, scr, screen[0], screen[1], vp.w, vp.h)
    ref.lib.ref_viewport_free(rv); ref.lib.ref_screen_free(scr); ref.lib.ref_scene_free(h)
    return px, z


@pytest.mark.parametrize("name,size", [("truck_4k_dof", (960, 540)), ("brainstem_4k_dof", (960, 540)), ("truck_4k_dof", (1920, 1080))])
def test_named_dof_workloads(ref_dofr, oracle, name, size):
    """the bench's DoF frames at a size the CPU finishes in seconds (the full 4K hashes are pinned by tools/make_golden.py)"""
    from swegl_b200.scene import Viewport
    scene, vps, screen, cfg = configs.build(name)
    vp = Viewport(0, 0, size[0], size[1], transparency_layers=0, post_mode=_abi.POST_DOF, focal_distance=5.0, focal_depth=5.0)
    vp.camera.apply(vps[0].pose)
    px, z = ref_frame(ref_dofr, scene, vp, size, vps[0].pose)
    o = oracle.render(scene, vp, screen_wh=size)
    assert (z.view(np.uint32) == o["z"].view(np.uint32)).all()
    assert (px == o["pixels"]).all(), f"{name}: DoF-R differs from the patched reference in {(px != o['pixels']).sum()} pixels"
    vp.post_mode = _abi.POST_NULL
    assert (oracle.render(scene, vp, screen_wh=size)["pixels"] != px).sum() > 1000      # the blur really did something


@pytest.mark.parametrize("seed", range(10))
@pytest.mark.parametrize("focal", [(5.0, 5.0), (4.0, 2.5), (3.5, 1.0), (6.0, 9.0), (0.5, 1.25)])
def test_fuzz_scenes(ref_dofr, oracle, seed, focal):
    """seeded triangle soups (near-plane clipping, slivers, all shader combinations) on a viewport at the screen origin --
    viewport_t::flatten reads the screen from row / column 0 whatever the viewport's offset (viewport.cpp:62), so only
    there is the DoF source the frame that was rendered"""
    from swegl_b200.scene import Viewport
    scene, vp0, screen, pose = configs.fuzz_case(seed)
    vp = Viewport(0, 0, screen[0] - seed, screen[1] - 2 * seed, light_mode=vp0.light_mode, tex_mode=vp0.tex_mode, transparency_layers=0,
                  post_mode=_abi.POST_DOF, focal_distance=focal[0], focal_depth=focal[1])
    vp.camera.apply(pose)
    px, z = ref_frame(ref_dofr, scene, vp, screen, pose)
    o = oracle.render(scene, vp, screen_wh=screen)
    assert (z.view(np.uint32) == o["z"].view(np.uint32)).all()
    assert (px == o["pixels"]).all()
