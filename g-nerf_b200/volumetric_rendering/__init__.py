from .renderer import ImportanceRenderer, PackedPlanes, pack_planes, pack_decoder, generate_planes  # noqa: F401
from .ray_sampler import RaySampler  # noqa: F401
from .ray_marcher import MipRayMarcher2  # noqa: F401
