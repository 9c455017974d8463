from nerf_downstream_b200.me.core import KernelGenerator, RegionType  # noqa: F401
