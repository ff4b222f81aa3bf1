"""Drop-in for ``audio_text_retrieval_criterion`` (criterions/audio_text_retrieval_loss.py) — identical to the
image twin except for the input keys and logging names ("[probed: diff]")."""
from ..fairseq_compat import register_criterion
from .image_text_retrieval_loss import ImageTextRetrievalCriterion


@register_criterion("audio_text_retrieval_criterion")
class AudioTextRetrievalCriterion(ImageTextRetrievalCriterion):
    a2b, b2a = "a2t_ncorrect", "t2a_ncorrect"

    def encode_other(self, model, ni):
        return model(src_audios=ni["src_audios"], audio_padding_masks=ni["audio_padding_masks"], encoder_type="audio")

    def compute_atc_loss(self, audio_logits, text_logits, audio_logits_all, text_logits_all, logit_scale_exp):
        return self.compute_itc_loss(audio_logits, text_logits, audio_logits_all, text_logits_all, logit_scale_exp)
