"""nrays_b200 — B200-native (sm_100a) implementation of the nrays render hot path.

Public surface mirrors the reference crate for this path (see scene.py); the pixels are produced
by hand-written CUDA behind the C-ABI of include/nrays_b200.h.  Importing the package does not
need a GPU; rendering does, and raises if the CUDA library is absent (no CPU fallback).
"""
from . import _abi
from .scene import (Ball, Capsule, Cone, Cuboid, Cylinder, FlatScene, Image, ImageData, Interpolation, Isometry3,
                    Light, Material, NormalMaterial, Overflow, PhongMaterial, Plane, Scene, SceneNode, Texture2d,
                    TriMesh, UVMaterial, camera_projection, make_camera, perspective3, render)

__all__ = [
    "Ball", "Capsule", "Cone", "Cuboid", "Cylinder", "FlatScene", "Image", "ImageData", "Interpolation",
    "Isometry3", "Light", "Material", "NormalMaterial", "Overflow", "PhongMaterial", "Plane", "Scene", "SceneNode",
    "Texture2d", "TriMesh", "UVMaterial", "camera_projection", "make_camera", "perspective3", "render",
]
