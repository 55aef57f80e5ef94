from .image_data import image_data  # noqa: F401
