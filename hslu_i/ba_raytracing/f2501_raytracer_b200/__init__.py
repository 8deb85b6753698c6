"""B200-native (sm_100a) render core for the per-pixel render loop of
kije/HSLU_I.BA_RAYTRACING.F2501_raytracer, behind the reference's Scene / Renderer API.

Importing the package does not need a GPU; every compute call goes through the C ABI of
librt_b200.so (include/rt_api.h) and raises if the library or a CUDA device is missing --
there is no CPU, PyTorch or oracle fallback on the product path.
"""
from . import _lib
from ._lib import RtError, load
from .output import FileOutput, get_feature_string, window_title
from .renderer import DeviceScene, ImageBuffer, MultiDeviceScene, RaytracerRenderer, RenderConfig
from .scene import FlatScene, Mesh, Scene, example_scene, material

__all__ = [
    "DeviceScene", "FileOutput", "FlatScene", "ImageBuffer", "Mesh", "MultiDeviceScene", "RaytracerRenderer", "RenderConfig", "RtError", "Scene",
    "example_scene", "get_feature_string", "load", "material", "window_title",
]
