"""Generate assets/*.scenepack from the reference's bundled glTF files (run in the build container,
where /root/reference exists).  The reference's OWN loader (src/data/gltf.cpp, compiled unmodified into
oracle/_ref) parses the files, so every loader quirk listed in SURVEY.md Appendix A.14 (uint16 indices,
swapped UV, matrix decomposition, meshes moved into the first referencing node ...) is baked into the
flattened arrays.  Embedded PNG/JPEG images are kept encoded (small) and decoded with PIL at load time;
the pack stores a digest of the decoded texels to detect decoder drift.

    python tools/make_scenepacks.py [/root/reference/resources]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle.binding import Ref  # noqa: E402
from swegl_b200.scene import Scene  # noqa: E402

SCENES = ["BoxTextured", "CesiumMilkTruck", "BrainStem", "BoxAnimated", "box"]


def main():
    res = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/resources"
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets")
    os.makedirs(out, exist_ok=True)
    ref = Ref()
    for name in SCENES:
        h = ref.load(os.path.join(res, name + ".glb"))
        s = ref.export(h, name)
        enc = list(ref.encoded_images) if len(ref.encoded_images) == len(s.textures) else None
        path = os.path.join(out, name + ".scenepack")
        s.save_pack(path, enc)
        back = Scene.load_pack(path)
        for a in Scene.ARRAYS:
            assert (getattr(back, a) == getattr(s, a)).all(), (name, a)
        for t0, t1 in zip(s.textures, back.textures):
            assert (t0 == t1).all(), (name, "texture")
        print(f"{name}: nodes={s.n_nodes} prims={s.n_primitives} verts={s.n_vertices} tris={s.n_triangles()} "
              f"textures={[t.shape for t in s.textures]} -> {os.path.getsize(path)} bytes")
        ref.lib.ref_scene_free(h)
    s = ref.procedural_scene()                      # built with the reference's own mesh builders, not a glTF
    s.save_pack(os.path.join(out, "procedural.scenepack"), None)
    print(f"procedural: prims={s.n_primitives} verts={s.n_vertices} tris={s.n_triangles()}")


if __name__ == "__main__":
    main()
