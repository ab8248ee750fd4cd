from .renderer import (ImportanceRenderer, PackedPlanes, pack_planes, pack_decoder, generate_planes,  # noqa: F401
                       project_onto_planes, sample_from_planes, sample_from_3dgrid)
from .ray_sampler import RaySampler  # noqa: F401
from .ray_marcher import MipRayMarcher2  # noqa: F401
