"""The example scenes `raytracing-potato_b200/scenes.py` BUILDS, compared with the constants of the reference's own
`src/example_scenes.rs` as frozen in tests/golden/scene_constants.json by tests/golden/make_scene_constants.py (a parser of the Rust
source, run where /root/reference exists). Oracle and GPU path consume the same scene description, so a typo in scenes.py would not
show up in any parity test; this is where it shows up. CPU only; loads the library for `rtp_camera_lookat`."""
import json
import math
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = json.load(open(os.path.join(HERE, "golden", "scene_constants.json")))["scenes"]
FOV = {"FRAC_PI_2": math.pi / 2.0, "FRAC_PI_4": math.pi / 4.0, "PI / 9.0": math.pi / 9.0}


def _kinds(rtp):
    A = rtp._abi
    return {
        "Scatter": {"None": A.SCATTER_NONE, "Lambert": A.SCATTER_LAMBERT, "Metal": A.SCATTER_METAL, "Dielectric": A.SCATTER_DIELECTRIC},
        "Absorb": {"BlackBody": A.ABSORB_BLACKBODY, "WhiteBody": A.ABSORB_WHITEBODY, "Albedo": A.ABSORB_ALBEDO, "AlbedoMap": A.ABSORB_ALBEDO_MAP},
        "Emit": {"None": A.EMIT_NONE, "DebugNormals": A.EMIT_DEBUG_NORMALS, "Color": A.EMIT_COLOR, "SkyGradient": A.EMIT_SKY_GRADIENT,
                 "SkySphere": A.EMIT_SKY_SPHERE},
        "Texture": {"Solid": A.TEXTURE_SOLID, "Image": A.TEXTURE_IMAGE, "Checker": A.TEXTURE_CHECKER, "Noise": A.TEXTURE_NOISE,
                    "Perlin": A.TEXTURE_PERLIN},
    }


def _check_component(rtp, got, want):
    kinds = _kinds(rtp)
    assert got.kind == kinds[want["family"]][want["kind"]], want
    if "param" in want:
        assert got.param == want["param"], want
    if "rgb" in want:
        assert np.array_equal(np.asarray(got.color, dtype=np.float64), np.asarray(want["rgb"])), want
    if "texture" in want:
        assert got.texture == want["texture"], want


def _check_camera(rtp, cam, want):
    assert cam.aspect_ratio == want["aspect_ratio"] and cam.focal_dist == want["focal_dist"] and cam.lens_radius == want["lens_radius"]
    assert cam.fov == FOV[want["fov"]]
    t = rtp.api.Transformation.lookat(want["eye"], want["target"], want["up"])
    assert np.array_equal(cam.transformation.position, np.asarray(want["eye"]))
    assert np.array_equal(cam.transformation.orientation, t.orientation)
    # utility.rs:172-177: z points from the target to the eye
    z = np.asarray(want["eye"]) - np.asarray(want["target"])
    assert np.allclose(cam.transformation.orientation[:, 2], z / np.linalg.norm(z), rtol=0, atol=1e-15)


def _check_tables(rtp, scene, want, n_materials=None):
    kinds = _kinds(rtp)
    textures = scene.scene_data.texture_table
    assert len(textures) == len(want["textures"])
    for got, w in zip(textures, want["textures"]):
        assert got.kind == kinds["Texture"][w["kind"]], w
        if w["kind"] == "Solid":
            assert np.array_equal(got.color, np.asarray(w["rgb"]))
        elif w["kind"] == "Checker":
            assert (got.odd, got.even) == (w["odd"], w["even"])
        elif w["kind"] in ("Perlin", "Noise"):
            assert got.seed == w["seed"]
        else:
            assert got.image is not None and got.image.ndim == 3 and got.image.shape[2] == 4
    materials = scene.scene_data.material_table
    assert len(materials) == (n_materials if n_materials is not None else len(want["materials"]))
    for got, w in zip(materials, want["materials"]):
        for part, wc in zip((got.scatter, got.absorb, got.emit), w):
            _check_component(rtp, part, wc)
    _check_component(rtp, scene.background, want["background"])


def _spheres(rtp, scene):
    h = scene.hittables
    return h[h["kind"] == rtp._abi.HITTABLE_SPHERE]


def _check_static_spheres(rtp, spheres, want):
    assert len(spheres) == len(want)
    for got, w in zip(spheres, want):
        assert np.array_equal(got["center"], np.asarray(w["center"])) and got["radius"] == w["radius"] and got["material"] == w["material"], w


@pytest.mark.parametrize("name", ["three_balls", "two_balls", "earth", "one_triangle", "bunny", "glass_bunny"])
def test_scene_equals_the_reference_source(rtp, name):
    want = FIXTURE[name]
    scene = getattr(rtp.scenes, name)()
    assert scene.root_kind == want["root"]
    _check_camera(rtp, scene.camera, want["camera"])
    _check_tables(rtp, scene, want)
    _check_static_spheres(rtp, _spheres(rtp, scene), want["spheres"])
    h = scene.hittables
    tri = h[h["kind"] == rtp._abi.HITTABLE_TRIANGLE]
    if want["whole_mesh_as_triangles"]:
        # hittable_list.extend(bunny.iter_triangles().map(...)); hittable_list.push(sphere): every triangle of mesh 0 in order, then the sphere
        mesh = scene.scene_data.mesh_table[0]
        assert len(scene.scene_data.mesh_table) == 1 and len(tri) == len(mesh.iter_triangles()) and len(tri) + 1 == len(h)
        assert np.array_equal(tri["triangle"], mesh.iter_triangles()) and not tri["mesh"].any()
        assert h[-1]["kind"] == rtp._abi.HITTABLE_SPHERE
        # the mesh is the OBJ file the reference names (assets.py bakes it from that file; the name is recorded with the arrays)
        assert os.path.basename(want["obj_files"][0]) in ("bunny.obj", "bunny_flat.obj")
        expect = rtp.assets.bunny() if want["obj_files"][0].endswith("bunny.obj") else rtp.assets.bunny_flat()
        assert np.array_equal(mesh.vertices["position"], expect.vertices["position"]) and np.array_equal(mesh.indices, expect.indices)
    else:
        assert [(int(t["triangle"]), int(t["mesh"])) for t in tri] == [(t["triangle"], t["mesh"]) for t in want["triangles"]]
    if "inline_mesh" in want:
        m, mesh = want["inline_mesh"], scene.scene_data.mesh_table[0]
        assert np.array_equal(mesh.vertices["position"], np.asarray(m["positions"]))
        n = np.asarray(m["normal_of"])
        assert np.allclose(mesh.vertices["normal"], n / math.sqrt(float(n @ n)), rtol=0, atol=1e-16)
        assert np.array_equal(mesh.vertices["uv"], np.tile(np.asarray(m["uv"]), (3, 1)))
        assert list(np.asarray(mesh.indices).ravel()) == m["indices"] and mesh.material == m["material"]
        # hittable order of the source: the triangle, then the ground sphere
        assert h[0]["kind"] == rtp._abi.HITTABLE_TRIANGLE and h[1]["kind"] == rtp._abi.HITTABLE_SPHERE


@pytest.mark.parametrize("name", ["more_balls", "more_balls_optimized"])
def test_more_balls_follows_the_reference_source(rtp, name):
    want, rnd = FIXTURE["more_balls"], FIXTURE["more_balls"]["random_part"]
    scene = getattr(rtp.scenes, name)()
    assert FIXTURE["more_balls_optimized"]["derived_from"] == "more_balls"
    assert scene.root_kind == FIXTURE[name]["root"]
    _check_camera(rtp, scene.camera, want["camera"])
    xs, zs = range(*rnd["x_range"]), [z for z in range(*rnd["z_range"]) if z != rnd["skipped_z"]]
    n_random = len(xs) * len(zs)
    _check_tables(rtp, scene, want, n_materials=len(want["materials"]) + n_random)
    spheres = _spheres(rtp, scene)
    assert len(spheres) == len(scene.hittables) == len(want["spheres"]) + n_random
    _check_static_spheres(rtp, spheres[:len(want["spheres"])], want["spheres"])
    # the random part: one sphere per (x, z) in the source's loop order, radius in ClosedRange(0.1, 0.3), resting on y = 0, inside its
    # cell by the source's offsets, with its own material appended in step
    rand = spheres[len(want["spheres"]):]
    lo, hi = rnd["radius_range"]
    assert rnd["offset_range"] == ["-0.5 + radius, 0.5 - radius"] * 2 and rnd["seed_byte"] == 249
    r = rand["radius"]
    assert (r >= lo).all() and (r <= hi).all() and np.array_equal(rand["center"][:, 1], r)
    cell_x = np.repeat(np.asarray(list(xs), dtype=np.float64), len(zs))
    cell_z = np.tile(np.asarray(zs, dtype=np.float64), len(xs))
    assert (np.abs(rand["center"][:, 0] - cell_x) <= 0.5 - r + 1e-12).all() and (np.abs(rand["center"][:, 2] - cell_z) <= 0.5 - r + 1e-12).all()
    assert np.array_equal(rand["material"], np.arange(len(want["materials"]), len(want["materials"]) + n_random))
    # Bernoulli(0.7) then Bernoulli(0.7): Lambert 0.7, Metal 0.21, glass 0.09 (refraction index from the source) - within 4 sigma
    A = rtp._abi
    kinds = np.array([m.scatter.kind for m in scene.scene_data.material_table[len(want["materials"]):]])
    p1, p2 = rnd["bernoulli"]
    for kind, p in ((A.SCATTER_LAMBERT, p1), (A.SCATTER_METAL, (1 - p1) * p2), (A.SCATTER_DIELECTRIC, (1 - p1) * (1 - p2))):
        assert abs((kinds == kind).mean() - p) < 4.0 * math.sqrt(p * (1 - p) / n_random)
    glass = [m for m in scene.scene_data.material_table[len(want["materials"]):] if m.scatter.kind == A.SCATTER_DIELECTRIC]
    assert all(m.scatter.param == rnd["glass_refraction_index"] and m.absorb.kind == A.ABSORB_WHITEBODY for m in glass)
    # draws per sphere in source order: radius, x offset, z offset, three albedo channels, Bernoulli (, Bernoulli (, fuzziness))
    assert rnd["draw_order"] == ["ClosedRange"] * 3 + [""] * 3 + ["Bernoulli", "Bernoulli", ""]
    rng = rtp.scenes.StdRngStream(bytes([rnd["seed_byte"]] * 32))
    first = spheres[len(want["spheres"])]
    radius = lo + rng.gen() * (hi - lo)
    assert first["radius"] == radius and first["center"][0] == float(xs[0]) + ((-0.5 + radius) + rng.gen() * ((0.5 - radius) - (-0.5 + radius)))


def test_fixture_covers_every_scene_function_of_the_reference():
    assert sorted(FIXTURE) == ["bunny", "earth", "glass_bunny", "more_balls", "more_balls_optimized", "one_triangle", "three_balls", "two_balls"]
