from nerf_downstream_b200.me.core import CoordinateManager, CoordinateMapKey, CoordinateMapType  # noqa: F401
