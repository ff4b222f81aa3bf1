from .recall import Recall  # noqa: F401
