"""CPU tests of the host-side mirror (scene XML subset, property parsing, flattening) and of the C ABI
library's symbol table (no compute calls without a GPU)."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import golden_util as gu
import mitsuba3dopplertof_b200 as dt
from mitsuba3dopplertof_b200 import _abi, runtime
from mitsuba3dopplertof_b200.transform import perspective_projection

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "dtof.h")).read()
    declared = set(re.findall(r"\b(dtof_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_abi.DTOF_SYMBOLS), declared ^ set(_abi.DTOF_SYMBOLS)
    if not os.path.exists(runtime.library_path()):
        runtime.build_library()
    lib = runtime.load_library()   # binds (and therefore resolves) every symbol
    assert lib.dtof_abi_version() == _abi.ABI_VERSION


def test_abi_struct_sizes_match_header():
    # sizes computed from include/dtof.h by the C compiler
    import subprocess, tempfile
    src = '#include "dtof.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",' \
          'sizeof(dtof_mesh),sizeof(dtof_instance),sizeof(dtof_bsdf),sizeof(dtof_emitter),sizeof(dtof_camera),' \
          'sizeof(dtof_film),sizeof(dtof_scene_desc),sizeof(dtof_params),sizeof(dtof_sample_record),sizeof(dtof_stats),' \
          'sizeof(dtof_pass_info));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")], check=True)
        sizes = list(map(int, subprocess.run([os.path.join(d, "s")], capture_output=True, text=True).stdout.split()))
    mine = [C.sizeof(t) for t in (_abi.Mesh, _abi.Instance, _abi.Bsdf, _abi.Emitter, _abi.Camera, _abi.Film,
                                  _abi.SceneDesc, _abi.Params, _abi.SampleRecord, _abi.Stats, _abi.PassInfo)]
    assert sizes == mine


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(dt.DTOFError):
        runtime.Context(0)
    scene = dt.load_file(os.path.join(gu.SCENES, "c1_example.xml"), resx=8, resy=8, spp=4)
    with pytest.raises(dt.DTOFError):
        scene.integrator.render(scene)


def test_example_scene_loads_unchanged():
    ref_xml = "/root/reference/configs_example/scene.xml"
    path = ref_xml if os.path.exists(ref_xml) else os.path.join(gu.SCENES, "c1_example.xml")
    sc = dt.load_file(path)
    integ = sc.integrator
    assert (integ.max_depth, integ.rr_depth, integ.path_correlation_depth) == (4, 5, 4)
    assert integ.time_sampling_method == "antithetic" and float(integ.antithetic_shift) == 0.5
    assert float(integ.hetero_frequency) == 1.0 and float(integ.sensor_phase_offset) == 0.0
    assert float(integ.w_s) == pytest.approx(30.0 + 1.0 / 0.0015 * 1e-6, rel=1e-6)
    assert sc.sensor.sampler.sample_count == 1024 and sc.sensor.sampler.time_correlate_number == 2
    flat = sc.flatten()
    assert (flat.n_triangles, flat.desc.n_instances, flat.desc.n_emitters) == (34, 3, 1)
    inst = flat.instances[1]
    assert inst.animated == 1 and inst.t0 == 0.0 and inst.t1 == np.float32(0.0015)
    assert inst.m1[11] - inst.m0[11] == pytest.approx(0.015, rel=1e-5)   # <translate z="0.015"/> on keyframe 1


def test_property_surface_and_errors():
    i = dt.DopplerToFPathIntegrator()
    assert (float(i.time), float(i.w_g), float(i.g_1), float(i.g_0), float(i.w_s)) == (np.float32(0.0015), 30.0, 0.5, 0.5, 30.0)
    assert i.time_sampling_method == "antithetic" and float(i.antithetic_shift) == 0.5
    assert i.use_stratified_sampling_for_each_interval and i.low_frequency_component_only
    assert (i.max_depth, i.rr_depth, i.path_correlation_depth) == (-1, 5, 0)
    assert float(dt.DopplerToFPathIntegrator(time_sampling_method="stratified").antithetic_shift) == 0.0
    j = dt.DopplerToFPathIntegrator(w_g=30.0, w_s=30.001)   # hetero_frequency derived from w_s - w_g
    assert float(j.hetero_frequency) == pytest.approx((np.float32(30.001) - np.float32(30.0)) * 1e6 * 0.0015, rel=1e-5)
    k = dt.DopplerToFPathIntegrator(hetero_offset=0.25)
    assert float(k.sensor_phase_offset) == pytest.approx(np.pi / 2, rel=1e-6)
    for bad in (dict(wave_function_type="sawtooth"), dict(time_sampling_method="periodic"), dict(rr_depth=0),
                dict(max_depth=-2), dict(not_a_property=1)):
        with pytest.raises(ValueError):
            dt.DopplerToFPathIntegrator(**bad)
    with pytest.raises(ValueError):   # README says sampler, the code says integrator (SURVEY.md finding 7)
        dt.load_string('<scene version="3.0.0"><integrator type="dopplertofpath"/><sensor type="perspective">'
                       '<sampler type="correlated"><boolean name="use_stratified_sampling_for_each_interval" value="true"/>'
                       '</sampler></sensor></scene>')
    s = dt.CorrelatedSampler(sample_count=8, time_correlate_number=4)
    assert s.path_correlate_number == 4
    with pytest.raises(ValueError):
        dt.DopplerToFPathIntegrator(time_sampling_method="antithetic_mirror").params(s)


def test_perspective_projection_matches_reference_headers():
    hv = json.load(open(os.path.join(gu.GOLDEN, "header_vectors.json")))
    for c in hv["perspective"]:
        c2s = perspective_projection(c["film"], c["crop"], c["offset"], c["fov"], c["near"], c["far"])
        s2c = c2s.inverse().matrix.astype(np.float32).reshape(16)
        ref = np.array(c["sample_to_camera"], np.float32)
        np.testing.assert_allclose(s2c, ref, rtol=3e-7, atol=1e-9)


def test_rectangle_winding_matches_frame_normal():
    from mitsuba3dopplertof_b200.scene import _flatten_shape, Shape
    for m in (np.diag([1.0, 1, 1, 1]), np.diag([1.0, -1, 1, 1]), np.diag([2.0, 3, -1, 1])):
        t = dt.Transform4.from_matrix(m)
        fm = _flatten_shape(Shape("rectangle"), t)
        p = fm.positions
        n_geo = np.cross(p[fm.faces[0][1]] - p[fm.faces[0][0]], p[fm.faces[0][2]] - p[fm.faces[0][0]])
        n_ref = t.astype(np.float32).transform_normal((0, 0, 1))
        assert np.dot(n_geo, n_ref) > 0


def test_mesh_loaders(tmp_path):
    from mitsuba3dopplertof_b200.meshio import load_mesh
    pos, faces, nrm, uv = load_mesh(os.path.join(gu.SCENES, "gem.ply"), face_normals=True)
    assert pos.shape == (6, 3) and faces.shape == (8, 3) and nrm is None and uv is None
    pos, faces, nrm, uv = load_mesh(os.path.join(gu.SCENES, "gem.ply"))
    assert nrm.shape == (6, 3) and np.allclose(np.linalg.norm(nrm, axis=1), 1, atol=1e-6)
    obj = tmp_path / "q.obj"
    obj.write_text("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nvn 0 0 1\n"
                   "f 1/1/1 2/2/1 3/3/1 4/4/1\n")
    pos, faces, nrm, uv = load_mesh(str(obj))
    assert faces.tolist() == [[0, 1, 2], [0, 2, 3]] and uv.shape == (4, 2) and nrm.shape == (4, 3)
    # flip_tex_coords defaults to true in the reference's obj loader (obj.cpp:151,266-267): v -> 1 - v
    assert uv.tolist() == [[0, 1], [1, 1], [1, 0], [0, 0]]
    assert load_mesh(str(obj), flip_tex_coords=False)[3].tolist() == [[0, 0], [1, 0], [1, 1], [0, 1]]


# ---- `.serialized` meshes (src/shapes/serialized.cpp) -----------------------------------------------------------------

def _ser_meshes():
    rng = np.random.default_rng(3)
    a = {"name": "a", "positions": rng.standard_normal((5, 3)), "faces": [[0, 1, 2], [2, 3, 4]],
         "normals": rng.standard_normal((5, 3)), "texcoords": rng.random((5, 2)), "colors": rng.random((5, 3))}
    b = {"name": "second mesh", "positions": rng.standard_normal((4, 3)), "faces": [[0, 1, 2], [0, 2, 3]]}
    return a, b


@pytest.mark.parametrize("version", [3, 4])
@pytest.mark.parametrize("double_precision", [False, True])
def test_serialized_round_trip(tmp_path, version, double_precision):
    from mitsuba3dopplertof_b200 import meshio
    a, b = _ser_meshes()
    path = str(tmp_path / "m.serialized")
    meshio.write_serialized(path, [a, b], version=version, double_precision=double_precision)
    for index, m in enumerate((a, b)):
        pos, faces, nrm, uv = meshio.load_serialized(path, index)
        assert pos.dtype == np.float32 and faces.dtype == np.uint32
        np.testing.assert_array_equal(pos, np.asarray(m["positions"]).astype(np.float32))   # narrowed like read_helper
        np.testing.assert_array_equal(faces, np.asarray(m["faces"], np.uint32))
        if "normals" in m:
            np.testing.assert_array_equal(nrm, np.asarray(m["normals"]).astype(np.float32))
            np.testing.assert_array_equal(uv, np.asarray(m["texcoords"]).astype(np.float32))
        else:
            assert nrm is None and uv is None
    # load_mesh: face_normals drops the file's normals, otherwise missing normals are computed (serialized.cpp:341-346,386-391)
    assert meshio.load_mesh(path, face_normals=True)[2] is None
    assert meshio.load_mesh(path, shape_index=1)[2].shape == (4, 3)


def test_serialized_errors_follow_the_reference(tmp_path):
    from mitsuba3dopplertof_b200 import meshio
    a, b = _ser_meshes()
    path = str(tmp_path / "m.serialized")
    meshio.write_serialized(path, [a, b])
    with pytest.raises(ValueError, match="shape index is out of range"):
        meshio.load_serialized(path, 2)
    with pytest.raises(ValueError, match="shape index must be nonnegative"):
        meshio.load_serialized(path, -1)
    with pytest.raises(ValueError, match="file not found"):
        meshio.load_serialized(str(tmp_path / "missing.serialized"))
    bad = tmp_path / "bad.serialized"
    bad.write_bytes(b"\x00\x00\x04\x00" + open(path, "rb").read()[4:])
    with pytest.raises(ValueError, match="invalid file format"):
        meshio.load_serialized(str(bad))
    bad.write_bytes(b"\x1c\x04\x07\x00" + open(path, "rb").read()[4:])
    with pytest.raises(ValueError, match="incompatible file version"):
        meshio.load_serialized(str(bad))


def test_serialized_scene_equals_the_ply_scene():
    """tests/scenes/c6_serialized.xml holds the gem of c5_slabroom.xml as sub-mesh 1 of a double-precision v4 file."""
    import golden_util as gu
    a = dt.load_file(os.path.join(gu.SCENES, "c6_serialized.xml")).flatten()
    b = dt.load_file(os.path.join(gu.SCENES, "c5_slabroom.xml")).flatten()
    assert a.n_triangles == b.n_triangles
    assert a.desc.n_meshes == b.desc.n_meshes
    for i in range(a.desc.n_meshes):
        ma, mb = a.desc.meshes[i], b.desc.meshes[i]
        assert (ma.n_vertices, ma.n_faces) == (mb.n_vertices, mb.n_faces)
        np.testing.assert_array_equal(np.ctypeslib.as_array(ma.positions, shape=(ma.n_vertices * 3,)),
                                      np.ctypeslib.as_array(mb.positions, shape=(mb.n_vertices * 3,)))
        np.testing.assert_array_equal(np.ctypeslib.as_array(ma.faces, shape=(ma.n_faces * 3,)),
                                      np.ctypeslib.as_array(mb.faces, shape=(mb.n_faces * 3,)))
        assert bool(ma.normals) == bool(mb.normals)


def test_conductor_bsdf_properties_follow_the_reference():
    """SmoothConductor ctor (src/bsdfs/conductor.cpp:213-230): material "none" = (eta 0, k 1); eta/k and a material
    together throw; unknown material names say so."""
    base = open(os.path.join(gu.SCENES, "c8_conductor.xml")).read()
    flat = dt.load_string(base, gu.SCENES, resx=16, resy=16, spp=4).flatten()
    kinds = {(b.kind, b.twosided): (tuple(b.eta), tuple(b.k), tuple(b.reflectance))
             for b in (flat.desc.bsdfs[i] for i in range(flat.desc.n_bsdfs))}
    mirror = kinds[(_abi.BSDF_CONDUCTOR, 1)]
    assert mirror == ((0.0, 0.0, 0.0), (1.0, 1.0, 1.0), (1.0, 1.0, 1.0))
    copper = kinds[(_abi.BSDF_CONDUCTOR, 0)]
    np.testing.assert_allclose(copper[0], (0.2004, 0.9240, 1.1022), rtol=1e-6)
    np.testing.assert_allclose(copper[2], (0.9, 0.95, 1.0), rtol=1e-6)
    with pytest.raises(ValueError, match="not in the table"):
        dt.load_string(base.replace('<bsdf type="conductor" />', '<bsdf type="conductor"><string name="material" value="Xx"/></bsdf>'),
                       gu.SCENES)
    with pytest.raises(ValueError, match="either"):
        dt.load_string(base.replace('<bsdf type="conductor" />',
                                    '<bsdf type="conductor"><string name="material" value="Au"/><rgb name="eta" value="1"/></bsdf>'), gu.SCENES)


def test_named_conductor_materials_come_from_the_reference_table():
    """conductor.cpp:221-229: `material` replaces eta / k. The table (conductor_ior.json, generated from the reference's
    own plugin by tests/golden/make_conductor_table.py) holds every material the reference itself can load."""
    from mitsuba3dopplertof_b200.xml_loader import conductor_ior
    eta, k = conductor_ior("Cu")
    np.testing.assert_allclose(eta, (0.201005474, 0.923749506, 1.10221541), rtol=1e-7)
    np.testing.assert_allclose(k, (3.91326213, 2.45304513, 2.14208984), rtol=1e-7)
    with pytest.raises(ValueError, match="not in the table"):
        conductor_ior("Unobtainium")
    flat = dt.load_file(os.path.join(gu.SCENES, "c13_named_metals.xml"), resx=16, resy=16, spp=4).flatten()
    by_kind = {b.kind: (tuple(b.eta), tuple(b.k)) for b in (flat.desc.bsdfs[i] for i in range(flat.desc.n_bsdfs))}
    np.testing.assert_allclose(by_kind[_abi.BSDF_CONDUCTOR][0], conductor_ior("Au")[0], rtol=1e-7)
    np.testing.assert_allclose(by_kind[_abi.BSDF_ROUGHCONDUCTOR][1], conductor_ior("Al")[1], rtol=1e-7)


def test_directional_emitter_properties_follow_the_reference():
    """DirectionalEmitter ctor (src/emitters/directional.cpp:65-91): `direction` is normalised and becomes the third column
    of to_world; `direction` and `to_world` together throw; unknown properties are reported."""
    base = open(os.path.join(gu.SCENES, "c16_directional.xml")).read()
    flat = dt.load_string(base, gu.SCENES, resx=16, resy=16, spp=4).flatten()
    assert flat.desc.n_emitters == 2
    e0, e1 = flat.desc.emitters[0], flat.desc.emitters[1]
    assert e0.kind == e1.kind == _abi.EMITTER_DIRECTIONAL
    want = np.array([0.35, -1.0, -0.6]) / np.linalg.norm([0.35, -1.0, -0.6])
    np.testing.assert_allclose(list(e0.position), want, rtol=3e-7)
    aim = np.array([0.05, 0.8, 0.0]) - np.array([0.1, 1.1, 6.8])
    np.testing.assert_allclose(list(e1.position), aim / np.linalg.norm(aim), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(list(e0.value), (6, 5.5, 5), rtol=1e-7)
    both = base.replace('<vector name="direction" value="0.35, -1.0, -0.6" />',
                        '<vector name="direction" value="0, -1, 0" /><transform name="to_world"><translate x="1" /></transform>')
    with pytest.raises(ValueError, match="Only one of the parameters"):
        dt.load_string(both, gu.SCENES)
    with pytest.raises(ValueError, match="unreferenced"):
        dt.load_string(base.replace('<vector name="direction" value="0.35, -1.0, -0.6" />', '<float name="radius" value="1" />'), gu.SCENES)


def test_film_options_that_would_change_the_image_are_refused():
    """sample_border (film.cpp:35, integrator.cpp:176-178) and a non-RGB pixel_format change what the reference writes;
    they are not built, and both hosts refuse them instead of ignoring them."""
    import subprocess
    base = open(os.path.join(gu.SCENES, "c1_example.xml")).read()
    fmt = '<string name="pixel_format" value="rgb" />'
    assert fmt in base
    border = base.replace(fmt, fmt + '<boolean name="sample_border" value="true" />')
    lum = base.replace(fmt, '<string name="pixel_format" value="luminance" />')
    off = base.replace(fmt, fmt + '<boolean name="sample_border" value="false" />')
    dt.load_string(off, gu.SCENES)
    with pytest.raises(ValueError, match="sample_border"):
        dt.load_string(border, gu.SCENES)
    with pytest.raises(ValueError, match="pixel_format"):
        dt.load_string(lum, gu.SCENES)
    cli = os.path.join(ROOT, "host", "dtof_render")
    assert subprocess.run(["make", "-C", os.path.join(ROOT, "host")], capture_output=True).returncode == 0
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        for xml, word in ((border, "sample_border"), (lum, "pixel_format")):
            path = os.path.join(tmp, "s.xml")
            open(path, "w").write(xml)
            r = subprocess.run([cli, "--dump-desc", os.path.join(tmp, "d.bin"), path], capture_output=True, text=True)
            assert r.returncode == 1 and word in r.stderr


def test_textured_and_unknown_property_values_are_refused_not_defaulted():
    """A <texture>, a named <ref> or a <spectrum filename=...> bound to a BSDF / emitter property is outside the subset.
    Silently dropping it would render with the plugin's default (0.5 reflectance, src/bsdfs/diffuse.cpp:83): both hosts
    refuse instead."""
    import subprocess
    import tempfile
    base = open(os.path.join(gu.SCENES, "c1_example.xml")).read()
    i = base.index('<bsdf type="diffuse"')
    j = base.index('>', i) + 1
    k = base.index('</bsdf>', j)
    variants = {
        "texture": base[:j] + '<texture name="reflectance" type="checkerboard"/>' + base[k:],
        "ref": base[:j] + '<ref name="reflectance" id="some_texture"/>' + base[k:],
        "spectrum": base[:j] + '<spectrum name="reflectance" filename="x.spd"/>' + base[k:],
        "unknown": base[:j] + '<volume name="reflectance" type="constvolume"/>' + base[k:],
    }
    cli = os.path.join(ROOT, "host", "dtof_render")
    assert subprocess.run(["make", "-C", os.path.join(ROOT, "host")], capture_output=True).returncode == 0
    with tempfile.TemporaryDirectory() as tmp:
        for what, xml in variants.items():
            with pytest.raises(ValueError, match="outside the supported subset"):
                dt.load_string(xml, gu.SCENES)
            path = os.path.join(tmp, "s.xml")
            open(path, "w").write(xml)
            r = subprocess.run([cli, "--dump-desc", os.path.join(tmp, "d.bin"), path], capture_output=True, text=True)
            assert r.returncode == 1 and "outside the supported subset" in r.stderr, (what, r.stderr)


def _independent_xml():
    base = open(os.path.join(gu.SCENES, "c1_example.xml")).read()
    i = base.index('<sampler type="correlated">')
    j = base.index('</sampler>', i) + len('</sampler>')
    return base[:i] + '<sampler type="independent"><integer name="sample_count" value="$spp" /></sampler>' + base[j:]


def test_dopplertofpath_with_the_independent_sampler_is_the_uniform_uncorrelated_stream():
    """With any sampler but `correlated`, the Doppler branch's next_1d_time / next_*_correlate calls hit the base-class
    defaults (include/mitsuba/render/sampler.h:131-144): one independent stream, whatever the integrator's
    time_sampling_method / path_correlation_depth say. That is bit for bit the correlated sampler's `rng` stream with
    uniform time sampling and no path correlation (src/samplers/correlated.cpp:92-97, 156-161)."""
    ind = dt.load_string(_independent_xml(), gu.SCENES, tsm="antithetic", pcd=4)
    assert ind.sensor.sampler.kind == "independent"
    cor = dt.load_file(os.path.join(gu.SCENES, "c1_example.xml"), tsm="uniform", pcd=0, tcn=1, pcn=1, strat="false")
    a = ind.integrator.params(ind.sensor.sampler, seed=3)
    b = cor.integrator.params(cor.sensor.sampler, seed=3)
    assert bytes(a) == bytes(b)
    assert bytes(ind.flatten().desc.camera) == bytes(cor.flatten().desc.camera)
    # the C++ host accepts the same scene
    import subprocess
    import tempfile
    cli = os.path.join(ROOT, "host", "dtof_render")
    assert subprocess.run(["make", "-C", os.path.join(ROOT, "host")], capture_output=True).returncode == 0
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "s.xml")
        open(path, "w").write(_independent_xml())
        r = subprocess.run([cli, "--dump-desc", os.path.join(tmp, "d.bin"), path], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_a_sensor_without_a_sampler_gets_the_reference_default():
    """src/render/sensor.cpp:47-48: no <sampler> child -> `independent`, 4 spp. Under dopplertofpath that is the uniform,
    uncorrelated stream (previous test), not this repo's correlated default."""
    import subprocess
    import tempfile
    base = open(os.path.join(gu.SCENES, "c1_example.xml")).read()
    i = base.index('<sampler type="correlated">')
    j = base.index('</sampler>', i) + len('</sampler>')
    bare = base[:i] + base[j:]
    no = dt.load_string(bare, gu.SCENES)
    explicit = dt.load_string(_independent_xml(), gu.SCENES, spp=4)
    assert no.sensor.sampler.kind == "independent" and no.sensor.sampler.sample_count == 4
    assert bytes(no.integrator.params(no.sensor.sampler, seed=1)) == bytes(explicit.integrator.params(explicit.sensor.sampler, seed=1))
    cli = os.path.join(ROOT, "host", "dtof_render")
    assert subprocess.run(["make", "-C", os.path.join(ROOT, "host")], capture_output=True).returncode == 0
    with tempfile.TemporaryDirectory() as tmp:
        outs = []
        for name, xml, extra in (("bare", bare, []), ("explicit", _independent_xml(), ["-Dspp=4"])):
            path = os.path.join(tmp, name + ".xml")
            open(path, "w").write(xml)
            out = os.path.join(tmp, name + ".bin")
            r = subprocess.run([cli] + extra + ["--dump-desc", out, path], capture_output=True, text=True)
            assert r.returncode == 0, r.stderr
            outs.append(open(out, "rb").read())
        assert outs[0] == outs[1]


def test_upload_cache_follows_scene_edits():
    """integrator.render() caches the upload per scene STATE: editing a shape, a material, the film or the sensor changes
    Scene.fingerprint(), so the stale resident copy is not rendered again (no GPU needed: a recording context stands in)."""
    from mitsuba3dopplertof_b200 import integrator as integ

    class FakeCtx:
        def __init__(self):
            self.uploads, self._flat = 0, None
        def upload(self, scene):
            self.uploads += 1
            self._flat = object()
            return self._flat

    scene = dt.load_file(os.path.join(gu.SCENES, "c1_example.xml"), resx=16, resy=16, spp=4)
    ctx = FakeCtx()
    a = integ._uploaded(ctx, scene)
    assert integ._uploaded(ctx, scene) is a and ctx.uploads == 1            # unchanged scene: one upload
    scene.sensor.film.width = 24
    b = integ._uploaded(ctx, scene)
    assert b is not a and ctx.uploads == 2                                   # another film size: uploaded again
    scene.shapes[0].bsdf = dt.scene.Bsdf(reflectance=(0.2, 0.3, 0.4)) if hasattr(dt, "scene") else scene.shapes[0].bsdf
    moved = [sh for sh in scene.shapes if sh.kind == "mesh" or sh.kind == "cube" or sh.kind == "rectangle"][0]
    moved.flip_normals = not moved.flip_normals
    assert integ._uploaded(ctx, scene) is not b and ctx.uploads == 3         # an edited shape: uploaded again
    other = FakeCtx()
    integ._uploaded(other, scene)
    assert other.uploads == 1                                                # another context: its own upload
