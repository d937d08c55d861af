from . import Conv2D, UpSampling2D, ZeroPadding2D, ZeroPadding3D  # noqa: F401
