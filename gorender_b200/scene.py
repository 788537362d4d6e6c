"""JSON scene manifests — host-side mirror of the reference's `scene.go`."""
from __future__ import annotations

import json
import logging
import os
from typing import Dict, List

import numpy as np

from .mesh import LoadMeshFile, Mesh, NewObject, Object
from .texture import LoadTextureFile, NewColorTexture
from .vecmath import vec3_to_radians

log = logging.getLogger("gorender_b200")


class Scene:
    """scene.go:32-54."""

    def __init__(self, Objects: List[Object] = None):
        self.Objects: List[Object] = list(Objects or [])

    def NumObjects(self) -> int:
        return len(self.Objects)

    def NumVertices(self) -> int:
        return sum(len(o.Mesh.Vertices) for o in self.Objects)

    def NumTriangles(self) -> int:
        return sum(len(o.Mesh.Faces) for o in self.Objects)


def LoadSceneFile(filename: str) -> Scene:
    """scene.go:56-133.  Manifest: {meshes:[{id,objFile,texture,textureScale}],
    objects:[{meshID,position,rotation(deg),scale}]} (scene.go:12-30)."""
    try:
        with open(filename, "r") as f:
            sceneData = json.load(f)
    except OSError as e:
        raise RuntimeError(f"failed to open level file: {e}")
    except json.JSONDecodeError as e:
        raise RuntimeError(f"failed to read scene manifest: {e}")

    defaultTexture = NewColorTexture((200, 200, 200, 255))  # scene.go:73-74
    rootDir = os.path.dirname(filename)
    meshes: Dict[str, Mesh] = {}
    objects: List[Object] = []

    for meshData in sceneData.get("meshes") or []:
        mid = meshData.get("id", "")
        try:
            loaded = LoadMeshFile(os.path.join(rootDir, meshData.get("objFile", "")), True)
        except Exception as e:
            raise RuntimeError(f"failed to load mesh '{mid}': {e}")
        mesh = loaded[0]
        if meshData.get("texture", "") != "":
            try:
                texture = LoadTextureFile(os.path.join(rootDir, meshData["texture"]))
            except Exception as e:
                raise RuntimeError(f"failed to load texture {mid}: {e}")
            if meshData.get("textureScale", 0) != 0:
                texture.SetScale(meshData["textureScale"])
            mesh.Faces.SetTexture(texture)
        else:
            mesh.Faces.SetTexture(defaultTexture)
        meshes[mid] = mesh

    for objData in sceneData.get("objects") or []:
        mesh = meshes.get(objData.get("meshID", ""))
        if mesh is None:
            raise RuntimeError(f"mesh id not found: {objData.get('meshID', '')}")
        scale = np.array(objData.get("scale", [0, 0, 0]), dtype=np.float32)
        if not scale.any():
            log.warning("object scale is zero: %s", objData.get("meshID", ""))
        obj = NewObject(mesh)
        obj.Scale = scale
        obj.Rotation = vec3_to_radians(np.array(objData.get("rotation", [0, 0, 0]), dtype=np.float32))
        obj.Translation = np.array(objData.get("position", [0, 0, 0]), dtype=np.float32)
        objects.append(obj)

    return Scene(objects)
