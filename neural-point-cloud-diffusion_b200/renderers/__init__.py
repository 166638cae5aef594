from .volume_renderer import VolumeRenderer
