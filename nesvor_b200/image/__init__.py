from .image import Image, Slice, Volume
