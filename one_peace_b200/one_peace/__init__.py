from .one_peace_base import OnePeaceBaseModel, ModelWrapper  # noqa: F401
from .one_peace_retrieval import OnePeaceRetrievalModel, OnePeaceRetrievalConfig  # noqa: F401
from .one_peace_pretrain import OnePeacePretrainModel, OnePeacePretrainConfig  # noqa: F401
