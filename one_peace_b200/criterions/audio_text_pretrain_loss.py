"""Drop-in for ``audio_text_pretrain_loss`` (criterions/audio_text_pretrain_loss.py:55-271): ATC + three DCL terms; the text
branch is a frozen teacher (:94-95, stage-2 pretraining).  Same constructor, model calls (:94-116), loss (:142-144) and
logging keys as the reference; arithmetic as in image_text_pretrain_loss.py here."""
import torch
import torch.distributed as dist

from ..fairseq_compat import register_criterion
from .image_text_pretrain_loss import _PretrainCriterionBase
from .image_text_retrieval_loss import gather_without_grad


@register_criterion("audio_text_pretrain_loss")
class AudioTextPretrainLossCriterion(_PretrainCriterionBase):
    loss_keys = ("atc_loss", "dcl_audio_loss", "dcl_al_text_loss", "dcl_al_audio_loss")
    acc_keys = ("a2t_ncorrect", "t2a_ncorrect")

    def __init__(self, task, dcl_audio_alpha=1.0, dcl_al_text_alpha=0.5, dcl_al_audio_alpha=0.5, dcl_logit_scale=2.5,
                 label_smoothing=0.0):
        super().__init__(task)
        self.dcl_audio_alpha = dcl_audio_alpha
        self.dcl_al_text_alpha = dcl_al_text_alpha
        self.dcl_al_audio_alpha = dcl_al_audio_alpha
        self.dcl_logit_scale = dcl_logit_scale
        self.label_smoothing = label_smoothing

    def forward(self, model, sample, reduce=True):
        """(loss, sample_size=1, logging_output) — audio_text_pretrain_loss.py:73-158."""
        ni = sample["net_input"]
        src_tokens, src_audios, apm = ni["src_tokens"], ni["src_audios"], ni["audio_padding_masks"]
        with torch.no_grad():
            text_logits, _ = model(src_tokens=src_tokens, encoder_type="text")
        audio_logits, _ = model(src_audios=src_audios, audio_padding_masks=apm, encoder_type="audio")
        text_all = gather_without_grad(text_logits) if dist.is_initialized() else text_logits.data
        audio_all = gather_without_grad(audio_logits) if dist.is_initialized() else audio_logits.data
        with torch.no_grad():
            teacher_al_text, teacher_al_audio = model(src_tokens=src_tokens, src_audios=src_audios, audio_padding_masks=apm,
                                                      encoder_type="al")
        _, _, student_audio = model(src_audios=src_audios, audio_preserve_ids=ni["audio_preserve_ids"],
                                    audio_padding_masks=apm, encoder_type="audio")
        student_al_text, _, student_al_audio = model(src_tokens=src_tokens, text_preserve_ids=ni["al_text_preserve_ids"],
                                                     src_audios=src_audios, audio_padding_masks=apm,
                                                     audio_preserve_ids=ni["al_audio_preserve_ids"], encoder_type="al")
        logit_scale_exp = model(return_logit_scale=True)
        text_pad = src_tokens.eq(1)
        audio_pad = apm[:, 1:]
        dcl_audio = self.compute_dcl_loss(student_audio, teacher_al_audio, ni["audio_mask_indices"], audio_pad)
        dcl_al_text = self.compute_dcl_loss(student_al_text, teacher_al_text, ni["al_text_mask_indices"], text_pad)
        dcl_al_audio = self.compute_dcl_loss(student_al_audio, teacher_al_audio, ni["al_audio_mask_indices"], audio_pad)
        atc, a2t_ok, t2a_ok = self.compute_atc_loss(audio_logits, text_logits, audio_all, text_all, logit_scale_exp)
        loss = atc + self.dcl_audio_alpha * dcl_audio + self.dcl_al_text_alpha * dcl_al_text + \
            self.dcl_al_audio_alpha * dcl_al_audio
        logging_output = {"loss": loss.data, "atc_loss": atc.data, "dcl_audio_loss": dcl_audio.data,
                          "dcl_al_text_loss": dcl_al_text.data, "dcl_al_audio_loss": dcl_al_audio.data,
                          "nsentences": sample["nsentences"], "sample_size": 1, "a2t_ncorrect": a2t_ok, "t2a_ncorrect": t2a_ok,
                          "logit_scale_exp": logit_scale_exp.data}
        return loss, 1, logging_output

    def compute_atc_loss(self, audio_logits, text_logits, audio_logits_all, text_logits_all, logit_scale_exp):
        return self._contrastive(audio_logits, text_logits, audio_logits_all, text_logits_all, logit_scale_exp)
