"""The C++ host (host/: XML reader, property surfaces, flattening; what a Mitsuba-side plugin / the dtof_render CLI
uses) and the Python mirror must hand IDENTICAL scene descriptions to the C ABI. CPU only: `dtof_render --dump-desc`
stops after flattening. The GPU test at the bottom renders through the CLI and compares with the Python path."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import golden_util as gu
import mitsuba3dopplertof_b200 as dt
from mitsuba3dopplertof_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "host", "dtof_render")


@pytest.fixture(scope="module")
def cli():
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "host")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return CLI


def _serialize_python(flat) -> bytes:
    out = [b"DTOFDESC1\n", np.uint32(flat.desc.n_meshes).tobytes()]
    for i in range(flat.desc.n_meshes):
        m = flat.meshes[i]
        nv, nf = m.n_vertices, m.n_faces
        out.append(np.array([nv, nf, bool(m.normals), bool(m.texcoords), m.bsdf, m.emitter & 0xFFFFFFFF, m.flip_normals, m.kind],
                            np.uint32).tobytes())
        out.append(np.array(list(m.rect_to_world), np.float32).tobytes())
        out.append(np.ctypeslib.as_array(m.positions, (nv * 3,)).tobytes())
        if m.normals:
            out.append(np.ctypeslib.as_array(m.normals, (nv * 3,)).tobytes())
        if m.texcoords:
            out.append(np.ctypeslib.as_array(m.texcoords, (nv * 2,)).tobytes())
        out.append(np.ctypeslib.as_array(m.faces, (nf * 3,)).tobytes())
    for n, arr, typ in ((flat.desc.n_instances, flat.instances, _abi.Instance), (flat.desc.n_bsdfs, flat.bsdfs, _abi.Bsdf),
                        (flat.desc.n_emitters, flat.emitters, _abi.Emitter)):
        out.append(np.uint32(n).tobytes())
        out.append(C.string_at(C.addressof(arr), n * C.sizeof(typ)))
    out.append(bytes(flat.desc.camera))
    out.append(bytes(flat.desc.film))
    return b"".join(out)


CASES = [
    ("c1_example.xml", {}),
    ("c1_example.xml", {"resx": 128, "resy": 96, "spp": 64, "tsm": "stratified", "wave": "triangular"}),
    ("c2_arealight.xml", {"resx": 512, "resy": 512}),
    ("c2b_two_emitters.xml", {}),
    ("c3_rotor.xml", {"pcn": 4}),
    ("c4_domino.xml", {"w_g": 150}),
    ("c5_slabroom.xml", {"tcn": 3, "pcn": 6}),
    ("c6_serialized.xml", {}),
    ("c7_constant.xml", {"max_depth": 5}),          # constant environment emitter + point light
    ("c8_conductor.xml", {"max_depth": 6}),         # smooth conductors (explicit eta / k, and the two-sided mirror)
    ("c9_dielectric.xml", {"max_depth": 8}),        # smooth dielectrics (named and numeric indices of refraction, tints)
    ("c10_thinglass.xml", {"max_depth": 8}),        # + a thin dielectric pane
    ("c11_plastic.xml", {"max_depth": 6}),          # smooth plastic (one- and two-sided, nonlinear, tinted coat)
    ("c12_roughconductor.xml", {"max_depth": 6}),   # rough conductors (GGX, anisotropic Beckmann)
    ("c13_named_metals.xml", {"max_depth": 6}),     # named conductor materials (Au, Al) from the generated table
    ("c14_spot.xml", {}),                            # spot light (lookat to_world, cutoff / beam angles)
    ("c15_roughdielectric.xml", {"max_depth": 8}),   # rough dielectrics (GGX, anisotropic Beckmann, tinted lobes)
    ("c16_directional.xml", {}),                     # directional lights (`direction` property and lookat to_world)
]


@pytest.mark.parametrize("scene,params", CASES)
def test_cpp_host_flattens_like_python_host(cli, tmp_path, scene, params):
    path = os.path.join(gu.SCENES, scene)
    dump = str(tmp_path / "desc.bin")
    args = [cli, "--dump-desc", dump] + [f"-D{k}={v}" for k, v in params.items()] + [path]
    r = subprocess.run(args, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = open(dump, "rb").read()
    want = _serialize_python(dt.load_file(path, **params).flatten())
    assert len(got) == len(want)
    if got != want:   # tolerate last-ulp differences of double-precision library routines, nothing else
        a, b = np.frombuffer(got[10:], np.uint8), np.frombuffer(want[10:], np.uint8)
        n = (len(a) // 4) * 4
        a, b = a[:n].view(np.uint32), b[:n].view(np.uint32)
        diff = np.nonzero(a != b)[0]
        fa, fb = a.view(np.float32)[diff], b.view(np.float32)[diff]
        # (+ 1e-15 absolute: an exactly-zero matrix entry of one host may be 1e-20 rounding noise in the other)
        assert np.all(np.abs(fa - fb) <= 2e-7 * np.maximum(np.abs(fb), 1e-30) + 1e-15), f"{len(diff)} words differ"
        assert len(diff) <= 8, f"{len(diff)} words differ between the C++ and the Python host"


def test_cpp_host_integrator_property_surface(cli, tmp_path):
    """Same names / defaults / errors as the reference constructors (dopplertofpath.cpp:19-57, integrator.cpp:54-100)."""
    base = open(os.path.join(gu.SCENES, "c1_example.xml")).read()

    def run(xml, *extra):
        p = tmp_path / "s.xml"
        p.write_text(xml)
        return subprocess.run([cli, "--dump-desc", str(tmp_path / "d.bin"), *extra, str(p)], capture_output=True, text=True)
    assert run(base).returncode == 0
    bad = base.replace('<float name="w_g" value="$w_g" />', '<float name="w_g" value="$w_g" /><float name="bogus" value="1" />')
    r = run(bad)
    assert r.returncode == 1 and "unreferenced property" in r.stderr
    r = run(base, "-Dwave=sawtooth")
    assert r.returncode == 1 and "wave_function_type" in r.stderr
    r = run(base, "-Dtsm=periodic")
    assert r.returncode == 1 and "time_sampling_method" in r.stderr
    r = run(base, "-Dnot_a_param=1")
    assert r.returncode == 1 and "Unused parameter" in r.stderr
    # use_stratified_sampling_for_each_interval belongs to the integrator, not the sampler (SURVEY.md finding 7)
    bad = base.replace('<integer name="time_correlate_number"', '<boolean name="use_stratified_sampling_for_each_interval" value="true" />'
                       '<integer name="time_correlate_number"')
    r = run(bad)
    assert r.returncode == 1 and "unreferenced property" in r.stderr


@pytest.mark.gpu
def test_cli_render_matches_python_path(cli, tmp_path):
    from mitsuba3dopplertof_b200 import runtime
    path = os.path.join(gu.SCENES, "c2_arealight.xml")
    out = str(tmp_path / "out.npy")
    r = subprocess.run([cli, "-Dresx=64", "-Dresy=48", "-Dspp=64", "-s", "3", "--raw", "-o", out, path], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Rendering finished." in r.stdout
    got = np.load(out)
    scene = dt.load_file(path, resx=64, resy=48, spp=64)
    ctx = runtime.Context(0)
    flat = ctx.upload(scene)
    want = ctx.render(flat, scene.integrator.params(scene.sensor.sampler, seed=3), develop=False)
    assert got.shape == want.shape == (48, 64, 4)
    # same inputs, same kernel: only the order of the float atomics differs
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()


def test_missing_vertex_normals_are_computed_after_to_world(cli, tmp_path):
    """A static obj / ply / serialized mesh without normals: the reference's loaders apply to_world to the positions and then
    run Mesh::recompute_vertex_normals on the WORLD-space vertices (obj.cpp:237,401-403, ply.cpp:288,435-437). Angle weights
    change under non-uniform scale, so transforming object-space normals is not the same. Both hosts, byte for byte."""
    from mitsuba3dopplertof_b200.meshio import vertex_normals
    base = open(os.path.join(gu.SCENES, "c5_slabroom.xml")).read()
    old = '<boolean name="face_normals" value="true" />'
    assert old in base
    xml = base.replace(old, "").replace('<scale value="0.22" />', '<scale x="0.4" y="0.1" z="0.22" />')
    for f in ("gem.ply",):
        os.symlink(os.path.join(gu.SCENES, f), str(tmp_path / f))
    path = str(tmp_path / "smooth.xml")
    open(path, "w").write(xml)
    scene = dt.load_file(path)
    flat = scene.flatten()
    gem = [i for i, sh in enumerate(scene.shapes) if sh.id == "Gem"]
    assert gem and scene.shapes[gem[0]].smooth_normals
    # the flattened mesh that has 6 vertices and 8 faces is the gem
    k = [i for i in range(flat.desc.n_meshes) if flat.meshes[i].n_vertices == 6 and flat.meshes[i].n_faces == 8]
    assert len(k) == 1
    m = flat.meshes[k[0]]
    pos = np.ctypeslib.as_array(m.positions, (18,)).reshape(6, 3).copy()
    nrm = np.ctypeslib.as_array(m.normals, (18,)).reshape(6, 3).copy()
    faces = np.ctypeslib.as_array(m.faces, (24,)).reshape(8, 3).copy()
    np.testing.assert_array_equal(nrm, vertex_normals(pos, faces))              # computed on the world-space positions
    sh = scene.shapes[gem[0]]
    obj_n = vertex_normals(sh.positions.astype(np.float32), sh.faces)
    t32 = sh.to_world.astype(np.float32)
    moved = np.stack([t32.transform_normal(n) for n in obj_n])
    moved /= np.linalg.norm(moved, axis=1, keepdims=True)
    assert np.abs(moved - nrm).max() > 1e-2                                     # ... which is not the transformed object-space result
    dump = str(tmp_path / "desc.bin")
    r = subprocess.run([cli, "--dump-desc", dump, path], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got, want = open(dump, "rb").read(), _serialize_python(flat)
    assert len(got) == len(want)
    a, b = np.frombuffer(got[10:len(got) // 4 * 4 + 2][: (len(got) - 10) // 4 * 4], np.uint8), np.frombuffer(want[10:][: (len(want) - 10) // 4 * 4], np.uint8)
    fa, fb = a.view(np.float32), b.view(np.float32)
    diff = np.nonzero(a.view(np.uint32) != b.view(np.uint32))[0]
    assert len(diff) <= 8 and np.all(np.abs(fa[diff] - fb[diff]) <= 2e-7 * np.maximum(np.abs(fb[diff]), 1e-30) + 1e-15)
