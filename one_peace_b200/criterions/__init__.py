from .image_text_retrieval_loss import ImageTextRetrievalCriterion  # noqa: F401
from .audio_text_retrieval_loss import AudioTextRetrievalCriterion  # noqa: F401
