"""B200-native (sm_100a) implementation of Mitsuba3DopplerToF's hot path: the `dopplertofpath`
integrator driven by the `correlated` sampler over motion-blurred scenes.

Host-side mirror of the reference's plugin surface (scene XML subset, integrator / sampler properties,
`render(scene, seed, spp)`) over the C ABI of `libdtof_b200.so` (include/dtof.h). Importing the package
needs neither the CUDA library nor a GPU; rendering does, and fails loudly without them.
"""
from .integrator import DopplerToFPathIntegrator, DTOFError, PathIntegrator, VelocityIntegrator
from .scene import (Bsdf, ConstantEmitter, CorrelatedSampler, Film, PerspectiveSensor, PointLight, Scene, Shape, SpotLight,
                    DirectionalLight, cube, mesh,
                    rectangle)
from .transform import AnimatedTransform, Transform4
from .xml_loader import load_file, load_string
from . import tof  # noqa: E402  (tutorial post-processing and drivers; needs the renderer only when called)

__all__ = [
    "DopplerToFPathIntegrator", "VelocityIntegrator", "PathIntegrator", "DTOFError", "Bsdf", "CorrelatedSampler", "Film", "PerspectiveSensor", "PointLight", "SpotLight", "DirectionalLight", "ConstantEmitter",
    "Scene", "Shape", "cube", "mesh", "rectangle", "AnimatedTransform", "Transform4", "load_file", "load_string", "tof",
]
