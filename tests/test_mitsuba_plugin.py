"""Drop-in boundary, end to end: the reference's OWN `mitsuba` executable (unmodified, scalar_rgb) renders a scene
through host/mitsuba_plugin/dopplertofpath_b200.cpp, i.e. XML parsing, plugin instantiation, scene graph, film and
file writing are the reference's code and only Integrator::render (src/render/integrator.cpp:104-347) is replaced by
the C ABI of libdtof_b200.so.

Needs a reference runtime in oracle/_ref/ (mitsuba, libmitsuba.so, plugins/, built out of tree from /root/reference by
the recipe in SURVEY.md Appendix C.1) with plugins/dopplertofpath_b200.so next to it (`make -C oracle/ref_harness`).
The runtime is git-ignored, so these tests skip on a checkout without it. Nothing here reads /root/reference.
"""
import os
import re
import subprocess

import numpy as np
import pytest

import mitsuba3dopplertof_b200 as dt
from mitsuba3dopplertof_b200 import runtime

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(ROOT, "oracle", "_ref")
EXE = os.path.join(REF, "mitsuba")
PLUGIN = os.path.join(REF, "plugins", "dopplertofpath_b200.so")
SCENES = os.path.join(HERE, "scenes")

needs_runtime = pytest.mark.skipif(not (os.path.exists(EXE) and os.path.exists(PLUGIN)),
                                   reason="no reference runtime + plugin in oracle/_ref (git-ignored build artefact)")


def _run(args, timeout=600, quiet=False):
    env = dict(os.environ, LD_LIBRARY_PATH=REF + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    if quiet:   # the scalar reference logs one warning per negative sample (hdrfilm.cpp:283-295): do not keep them
        return subprocess.run([EXE, "-m", "scalar_rgb"] + args, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=env,
                              timeout=timeout)
    return subprocess.run([EXE, "-m", "scalar_rgb"] + args, capture_output=True, text=True, env=env, timeout=timeout)


def _runnable():
    """The runtime was compiled with -march=native in the build container; skip where that ISA is missing."""
    try:
        return _run(["--help"], timeout=60).returncode == 0
    except Exception:   # noqa: BLE001
        return False


def _read_pfm(path):
    with open(path, "rb") as f:
        kind = f.readline().strip()
        w, h = map(int, f.readline().split())
        scale = float(f.readline())
        data = np.frombuffer(f.read(), "<f4" if scale < 0 else ">f4")
    ch = 3 if kind == b"PF" else 1
    return data.reshape(h, w, ch)[::-1].astype(np.float32)   # PFM stores the bottom row first


def _plugin_scene(tmp_path, name, integrator="dopplertofpath_b200"):
    xml = open(os.path.join(SCENES, name)).read()
    xml = xml.replace('<integrator type="dopplertofpath">', f'<integrator type="{integrator}">')
    xml = xml.replace('<string name="file_format" value="openexr" />',
                      '<string name="file_format" value="pfm" />\n<string name="component_format" value="float32" />')
    path = os.path.join(str(tmp_path), name)
    open(path, "w").write(xml)
    for f in os.listdir(SCENES):
        if f.endswith((".ply", ".serialized")) and not os.path.exists(os.path.join(str(tmp_path), f)):
            os.symlink(os.path.join(SCENES, f), os.path.join(str(tmp_path), f))
    return path


@needs_runtime
def test_plugin_exports_the_mitsuba_plugin_abi():
    """PluginManager reads `plugin_name` / `plugin_descr` after dlopen (src/core/plugin.cpp:21-49)."""
    out = subprocess.run(["nm", "-D", "--defined-only", PLUGIN], capture_output=True, text=True).stdout
    assert re.search(r"\bplugin_name\b", out) and re.search(r"\bplugin_descr\b", out)
    # it binds the product through the C ABI only
    und = subprocess.run(["nm", "-D", "--undefined-only", PLUGIN], capture_output=True, text=True).stdout
    for sym in ("dtof_create", "dtof_upload_scene", "dtof_render_accumulate", "dtof_read_film", "dtof_last_error", "dtof_destroy"):
        assert re.search(rf"\b{sym}\b", und), sym


@needs_runtime
def test_plugin_fails_loudly_without_a_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    if not _runnable():
        pytest.skip("reference runtime cannot execute on this CPU")
    r = _run(["-Dspp=4", "-Dresx=16", "-Dresy=16", "-o", os.path.join(str(tmp_path), "o.pfm"), _plugin_scene(tmp_path, "c1_example.xml")])
    assert r.returncode != 0
    assert "no usable CUDA device (there is no CPU fallback)" in (r.stdout + r.stderr)


@pytest.mark.gpu
@needs_runtime
@pytest.mark.parametrize("name,defs", [
    ("c1_example.xml", {"resx": 96, "resy": 96, "spp": 128}),
    ("c2_arealight.xml", {"resx": 96, "resy": 64, "spp": 64}),
    ("c4_domino.xml", {"resx": 128, "resy": 96, "spp": 64, "wave": "trapezoidal", "tsm": "antithetic_mirror", "shift": 0.0, "w_g": 150}),
    ("c5_slabroom.xml", {"resx": 64, "resy": 64, "spp": 32}),
    ("c6_serialized.xml", {"resx": 64, "resy": 64, "spp": 32}),                          # reference's SerializedMesh loader
    ("c7_constant.xml", {"resx": 64, "resy": 48, "spp": 32, "hetero_frequency": 0.0}),   # constant environment emitter
    ("c8_conductor.xml", {"resx": 64, "resy": 64, "spp": 32, "hetero_frequency": 0.0, "max_depth": 6}),   # conductors
    ("c9_dielectric.xml", {"resx": 64, "resy": 64, "spp": 32, "hetero_frequency": 0.0, "max_depth": 8}),   # dielectrics
    ("c10_thinglass.xml", {"resx": 64, "resy": 64, "spp": 32, "hetero_frequency": 0.0, "max_depth": 8}),   # + thin pane
    ("c11_plastic.xml", {"resx": 64, "resy": 64, "spp": 32, "hetero_frequency": 0.0, "max_depth": 6}),     # plastic
    ("c12_roughconductor.xml", {"resx": 64, "resy": 64, "spp": 32, "hetero_frequency": 0.0, "max_depth": 6}),   # microfacets
    # named materials: the reference converts its measured spectra, the Python host looks the generated table up
    ("c13_named_metals.xml", {"resx": 64, "resy": 64, "spp": 32, "hetero_frequency": 0.0, "max_depth": 6}),
    ("c14_spot.xml", {"resx": 64, "resy": 64, "spp": 32, "hetero_frequency": 0.0}),      # spot light
    ("c15_roughdielectric.xml", {"resx": 64, "resy": 64, "spp": 32, "hetero_frequency": 0.0, "max_depth": 8}),   # frosted glass
    ("c16_directional.xml", {"resx": 64, "resy": 64, "spp": 32, "hetero_frequency": 0.0}),     # distant lights
])
def test_reference_cli_renders_through_the_plugin(tmp_path, name, defs):
    """`mitsuba -m scalar_rgb scene.xml` with integrator dopplertofpath_b200 == the Python host's render of the same
    scene: the two hosts flatten independently (reference scene graph vs. this repository's XML reader) and must hand
    the kernels the same arrays. Film sums are unordered float atomics, hence the small tolerance; the two hosts list
    shapes in different orders, so the BVHs differ and a sample that grazes a shared triangle edge may resolve to the
    other triangle (measured: 1 sample in 786 k on the domino scene) -- at most 0.1 % of the pixels may be off."""
    if not _runnable():
        pytest.skip("reference runtime cannot execute on this CPU")
    out = os.path.join(str(tmp_path), "out.pfm")
    r = _run([f"-D{k}={v}" for k, v in defs.items()] + ["-o", out, _plugin_scene(tmp_path, name)])
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    assert "Rendering finished." in r.stdout + r.stderr
    img = _read_pfm(out)
    scene = dt.load_file(os.path.join(SCENES, name), **defs)
    ref = scene.integrator.render(scene, seed=0)
    assert img.shape == ref.shape
    scale = np.abs(ref).max()
    assert scale > 0
    err = np.abs(img - ref).max(axis=2)
    bad = int((err > 2e-4 * scale).sum())
    assert bad <= 1e-3 * err.size, f"plugin render differs from the Python host in {bad} pixels (max {err.max():.3e}, scale {scale:.3e})"
    assert np.median(err) <= 2e-6 * scale


@pytest.mark.gpu
@needs_runtime
@pytest.mark.parametrize("rfilter", [
    '<rfilter type="mitchell" />',
    '<rfilter type="mitchell"><float name="B" value="0.2" /><float name="C" value="0.6" /></rfilter>',
    '<rfilter type="catmullrom" />',
    '<rfilter type="lanczos"><integer name="lobes" value="2" /></rfilter>',
    '<rfilter type="gaussian"><float name="stddev" value="0.8" /></rfilter>',
])
def test_reference_cli_filters_through_the_plugin(tmp_path, rfilter):
    """The film's reconstruction filter as the reference's scene graph holds it (class name, radius, Mitchell's B / C)
    reaches the kernels like the Python host's reading of the same XML."""
    if not _runnable():
        pytest.skip("reference runtime cannot execute on this CPU")
    src = open(os.path.join(SCENES, "c1_example.xml")).read()
    assert '<rfilter type="tent" />' in src
    plain = os.path.join(str(tmp_path), "plain.xml")
    open(plain, "w").write(src.replace('<rfilter type="tent" />', rfilter))
    scene_path = _plugin_scene(tmp_path, "c1_example.xml")
    xml = open(scene_path).read().replace('<rfilter type="tent" />', rfilter)
    open(scene_path, "w").write(xml)
    defs = {"resx": 48, "resy": 48, "spp": 32}
    out = os.path.join(str(tmp_path), "out.pfm")
    r = _run([f"-D{k}={v}" for k, v in defs.items()] + ["-o", out, scene_path])
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    img = _read_pfm(out)
    scene = dt.load_file(plain, **defs)
    ref = scene.integrator.render(scene, seed=0)
    assert img.shape == ref.shape
    scale = np.abs(ref).max()
    err = np.abs(img - ref).max(axis=2)
    assert int((err > 2e-4 * scale).sum()) <= 1e-3 * err.size, f"max {err.max():.3e}, scale {scale:.3e}"


@pytest.mark.gpu
@needs_runtime
def test_reference_cli_independent_sampler_through_the_plugin(tmp_path):
    """`independent` sampler + Doppler integrator: the plugin (reading the reference's IndependentSampler) and the Python
    host must choose the same stream -- the correlated sampler's independent stream, uniform in time, uncorrelated."""
    if not _runnable():
        pytest.skip("reference runtime cannot execute on this CPU")
    def swap(xml):
        i = xml.index('<sampler type="correlated">')
        j = xml.index('</sampler>', i) + len('</sampler>')
        return xml[:i] + '<sampler type="independent"><integer name="sample_count" value="$spp" /></sampler>' + xml[j:]
    scene_path = _plugin_scene(tmp_path, "c1_example.xml")
    swapped = swap(open(scene_path).read())
    open(scene_path, "w").write(swapped)
    plain = os.path.join(str(tmp_path), "plain.xml")
    open(plain, "w").write(swap(open(os.path.join(SCENES, "c1_example.xml")).read()))
    defs = {"resx": 48, "resy": 48, "spp": 32, "tcn": 2, "pcn": 2}
    out = os.path.join(str(tmp_path), "out.pfm")
    r = _run([f"-D{k}={v}" for k, v in defs.items() if k not in ("tcn", "pcn")] + ["-o", out, scene_path])
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    img = _read_pfm(out)
    scene = dt.load_file(plain, **{k: v for k, v in defs.items() if k not in ("tcn", "pcn")})
    ref = scene.integrator.render(scene, seed=0)
    same = dt.load_file(os.path.join(SCENES, "c1_example.xml"), tsm="uniform", pcd=0, tcn=1, pcn=1, strat="false",
                        **{k: v for k, v in defs.items() if k not in ("tcn", "pcn")})
    ref2 = same.integrator.render(same, seed=0)
    scale = np.abs(ref).max()
    assert np.abs(ref - ref2).max() <= 2e-4 * scale
    err = np.abs(img - ref).max(axis=2)
    assert int((err > 2e-4 * scale).sum()) <= 1e-3 * err.size, f"max {err.max():.3e}, scale {scale:.3e}"


@pytest.mark.gpu
@needs_runtime
def test_plugin_image_agrees_with_the_reference_integrator(tmp_path):
    """Same executable, same scene file, integrator `dopplertofpath` (the reference's CPU code) vs `dopplertofpath_b200`.
    The scalar reference is not stream-identical (per-pixel seeding, no pair correlation), so the check
    is statistical, on the homodyne image where the scalar estimator's variance is low: the mean of the B200 image
    must lie within the scalar render's own Monte-Carlo error of its mean, region by region."""
    if not _runnable():
        pytest.skip("reference runtime cannot execute on this CPU")
    defs = {"resx": 64, "resy": 64, "spp": 256, "hetero_frequency": 0.0}
    imgs = {}
    for integ in ("dopplertofpath", "dopplertofpath_b200"):
        out = os.path.join(str(tmp_path), integ + ".pfm")
        r = _run([f"-D{k}={v}" for k, v in defs.items()] + ["-o", out, _plugin_scene(tmp_path, "c1_example.xml", integ)],
                 quiet=True)
        assert r.returncode == 0
        imgs[integ] = _read_pfm(out).astype(np.float64)
    a, b = imgs["dopplertofpath"], imgs["dopplertofpath_b200"]
    # 4 x 4 regions of 16 x 16 pixels: region means, error bar from the scalar image's pixel-to-pixel spread
    for ry in range(4):
        for rx in range(4):
            sa = a[16 * ry:16 * ry + 16, 16 * rx:16 * rx + 16, 1]
            sb = b[16 * ry:16 * ry + 16, 16 * rx:16 * rx + 16, 1]
            sem = np.sqrt((sa.var() + sb.var()) / sa.size) + 1e-7
            assert abs(sa.mean() - sb.mean()) < 6 * sem + 0.02 * abs(sa.mean()), (ry, rx, sa.mean(), sb.mean(), sem)


# ---- the drop-in NAME: an untouched scene file through the reference's executable -------------------------------------
DROPIN = os.path.join(REF, "dropin")
DROPIN_EXE = os.path.join(DROPIN, "mitsuba")
DROPIN_PLUGIN = os.path.join(DROPIN, "plugins", "dopplertofpath.so")
REFERENCE_SCENE = os.path.join(HERE, "golden", "reference_scene.xml")   # configs_example/scene.xml, byte for byte
needs_dropin = pytest.mark.skipif(not (os.path.exists(DROPIN_EXE) and os.path.exists(DROPIN_PLUGIN)),
                                  reason="no drop-in view of the reference runtime in oracle/_ref/dropin (git-ignored build artefact)")


def _run_dropin(args, timeout=600, env_extra=None):
    """The reference's executable started from the drop-in view: libmitsuba.so is found there, so plugins/<type>.so
    resolves there too (src/mitsuba/mitsuba.cpp:315-318) and plugins/dopplertofpath.so is the product's plugin."""
    env = dict(os.environ, LD_LIBRARY_PATH=DROPIN + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    env.update(env_extra or {})
    return subprocess.run([DROPIN_EXE, "-m", "scalar_rgb"] + args, capture_output=True, text=True, env=env, timeout=timeout)


def _read_exr(path):
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    import cv2
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    assert img is not None, path
    return np.ascontiguousarray(img[..., :3][..., ::-1]).astype(np.float32)   # BGR -> RGB


def test_reference_scene_fixture_is_the_reference_file():
    """tests/golden/reference_scene.xml is configs_example/scene.xml byte for byte (checked where the reference is present)."""
    ref = "/root/reference/configs_example/scene.xml"
    if not os.path.exists(ref):
        pytest.skip("no reference checkout here")
    assert open(ref, "rb").read() == open(REFERENCE_SCENE, "rb").read()
    assert b'<integrator type="dopplertofpath">' in open(REFERENCE_SCENE, "rb").read()


@needs_dropin
def test_dropin_view_shadows_only_the_integrator():
    names = set(os.listdir(os.path.join(DROPIN, "plugins")))
    assert "dopplertofpath.so" in names and "correlated.so" in names and "hdrfilm.so" in names
    assert not os.path.islink(DROPIN_PLUGIN)                                   # the product's plugin, a real file
    assert os.path.islink(os.path.join(DROPIN, "plugins", "correlated.so"))    # everything else is the reference's
    und = subprocess.run(["nm", "-D", "--undefined-only", DROPIN_PLUGIN], capture_output=True, text=True).stdout
    for sym in ("dtof_create", "dtof_create_multi", "dtof_render_accumulate", "dtof_read_film"):
        assert re.search(rf"\b{sym}\b", und), sym


@pytest.mark.gpu
@needs_dropin
def test_untouched_reference_scene_renders_through_the_dropin_plugin(tmp_path):
    """`mitsuba scene.xml` with configs_example/scene.xml UNCHANGED (integrator type "dopplertofpath", 256 x 256 @ 1024
    spp, OpenEXR output): XML parsing, the correlated sampler, film and file writing are the reference's code, the
    integrator is the B200 plugin. The image is compared with the reference's own committed render of this file
    (configs_example/scene.exr, tests/golden/scene_exr.npy)."""
    if not _runnable():
        pytest.skip("reference runtime cannot execute on this CPU")
    scene = os.path.join(str(tmp_path), "scene.xml")
    with open(scene, "wb") as f:
        f.write(open(REFERENCE_SCENE, "rb").read())
    out = os.path.join(str(tmp_path), "scene.exr")
    r = _run_dropin(["-o", out, scene])
    log = r.stdout + r.stderr
    assert r.returncode == 0, log[-2000:]
    assert "dopplertofpath on B200" in log and "Rendering finished." in log
    img = _read_exr(out)
    gold = np.load(os.path.join(HERE, "golden", "scene_exr.npy")).astype(np.float32)
    assert img.shape == gold.shape == (256, 256, 3)
    rel_mse = float(((img - gold) ** 2).mean() / (gold ** 2).mean())
    assert rel_mse < 5e-3, f"relative MSE vs configs_example/scene.exr = {rel_mse:.3e}"


@pytest.mark.gpu
@needs_dropin
def test_dropin_plugin_honours_timeout(tmp_path):
    """`timeout` (include/mitsuba/render/integrator.h:106-108) is checked between the chunks the plugin submits: a
    2 x 2048-spp render of the domino scene (16 chunks of 2^28 lanes) with a 50 ms budget stops early and still writes a
    (partially accumulated) image."""
    if not _runnable():
        pytest.skip("reference runtime cannot execute on this CPU")
    xml = open(os.path.join(SCENES, "c4_domino.xml")).read()
    xml = xml.replace('<integrator type="dopplertofpath">', '<integrator type="dopplertofpath">\n<float name="timeout" value="0.05" />')
    xml = xml.replace('<string name="file_format" value="openexr" />',
                      '<string name="file_format" value="pfm" />\n<string name="component_format" value="float32" />')
    scene = os.path.join(str(tmp_path), "c4.xml")
    open(scene, "w").write(xml)
    out = os.path.join(str(tmp_path), "out.pfm")
    defs = {"resx": 1024, "resy": 1024, "spp": 4096, "wave": "trapezoidal", "tsm": "antithetic_mirror", "shift": 0.0, "w_g": 150}
    r = _run_dropin([f"-D{k}={v}" for k, v in defs.items()] + ["-o", out, scene])
    log = r.stdout + r.stderr
    assert r.returncode == 0, log[-2000:]
    assert "Rendering stopped early" in log
    img = _read_pfm(out)
    assert np.isfinite(img).all()
    # the first chunks cover the top rows; the bottom of the frame was never reached
    assert np.abs(img[:32]).max() > 0 and np.abs(img[-32:]).max() == 0


@pytest.mark.gpu
@needs_dropin
def test_dropin_plugin_on_all_gpus_equals_one_gpu(tmp_path):
    """DTOF_DEVICES=all: the reference's single `mitsuba` process renders the untouched scene on every GPU of the box
    (one context, sharded behind the C ABI); the image equals the one-GPU image to float summation order."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    if not _runnable():
        pytest.skip("reference runtime cannot execute on this CPU")
    scene = os.path.join(str(tmp_path), "scene.xml")
    with open(scene, "wb") as f:
        f.write(open(REFERENCE_SCENE, "rb").read())
    imgs = []
    for devs in ("0", "all"):
        out = os.path.join(str(tmp_path), f"scene_{devs}.exr")
        r = _run_dropin(["-o", out, scene], env_extra={"DTOF_DEVICES": devs})    # the file is untouched: no -D parameters
        log = r.stdout + r.stderr
        assert r.returncode == 0, log[-2000:]
        assert f"on {torch.cuda.device_count() if devs == 'all' else 1} device(s)" in log
        imgs.append(_read_exr(out))
    scale = np.abs(imgs[0]).max()
    assert np.abs(imgs[0] - imgs[1]).max() <= 2e-3 * scale    # fp16 EXR: half-precision rounding of each image
