"""gorender_b200 — B200-native implementation of gorender's per-frame hot path.

The public names mirror the reference's Go API.  Importing the
package does not load the native library; the first `Device` / `FrameBuffer`
does, and fails loudly if it has not been built.
"""
from .vecmath import (NewIdentityMatrix, NewScaleMatrix, NewTranslationMatrix, NewRotationXMatrix,
                      NewRotationYMatrix, NewRotationZMatrix, NewRotationMatrix, NewWorldMatrix,
                      NewPerspectiveMatrix, NewScreenMatrix, NewLookAtMatrix, NewViewMatrix, Multiply, Transpose)
from .texture import (Texture, NewColorTexture, NewImageTexture, LoadTextureFile, TextureTypeSolidColor,
                      TextureTypeImage, TextureTypeImageFast)
from .mesh import FaceArray, Mesh, Object, NewMesh, NewObject, LoadMeshFile, boundingBox
from .obj import LoadObjFile, LoadObjFileNative
from .scene import Scene, LoadSceneFile
from .renderer import Camera, Device, FrameBuffer, Renderer, NewFrameBuffer, NewRenderer, default_device

__all__ = [n for n in dir() if not n.startswith("_")]
