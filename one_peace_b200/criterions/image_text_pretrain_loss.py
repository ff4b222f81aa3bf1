"""Drop-in for ``image_text_pretrain_loss`` (criterions/image_text_pretrain_loss.py:56-279): ITC + four DCL terms.

Same constructor arguments (``dcl_*_alpha``, ``dcl_logit_scale``, ``label_smoothing``), same six model calls in the same
order (:98-121), same loss combination (:145-147) and logging keys (:149-161).  ``compute_itc_loss`` is the InfoNCE head of
``image_text_retrieval_loss`` (tcgen05 GEMM epilogues, csrc/infonce.cu); ``compute_dcl_loss`` (:187-208) is ONE direction of
the same tiled similarity + log-softmax kernels with a row selection in front (``autograd_general.DclLossFn``): masked
student rows against every non-padded teacher row of the local batch, fp32 L2-normalisation, scale ``dcl_logit_scale``,
label-smoothed NLL — the (n_masked x n_teacher) logits never reach HBM in fp32.
"""
import torch
import torch.distributed as dist

from ..autograd_general import DclLossFn, dcl_indices
from ..fairseq_compat import FairseqCriterion, metrics, register_criterion
from .image_text_retrieval_loss import gather_without_grad, itc_loss


def dcl_loss(student_features, teacher_features, mask_indices, padding_masks=None, dcl_logit_scale=2.5, label_smoothing=0.0):
    """compute_dcl_loss (image_text_pretrain_loss.py:187-208).  student / teacher (B,S,d); mask_indices bool (B,S);
    padding_masks bool (B,S-1) or None.  Gradient flows to the student only (:189)."""
    B, S, d = student_features.shape
    stu_idx, tea_idx = dcl_indices(mask_indices, padding_masks)
    if stu_idx.numel() == 0:
        raise RuntimeError("compute_dcl_loss: no masked token in the batch")
    return DclLossFn.apply(student_features.reshape(B * S, d), teacher_features.detach().reshape(B * S, d), stu_idx, tea_idx,
                           float(dcl_logit_scale), float(label_smoothing)).to(student_features.dtype)


class _PretrainCriterionBase(FairseqCriterion):
    """Shared pieces of the two pretraining criteria (the reference duplicates them per file)."""
    loss_keys = ()
    acc_keys = ()

    def compute_dcl_loss(self, student_features, teacher_features, mask_indices, padding_masks=None):
        return dcl_loss(student_features, teacher_features, mask_indices, padding_masks, self.dcl_logit_scale, self.label_smoothing)

    def _contrastive(self, a_logits, text_logits, a_all, text_all, logit_scale_exp):
        rank = dist.get_rank() if dist.is_initialized() else 0
        return itc_loss(a_logits, text_logits, a_all, text_all, logit_scale_exp, rank, 0.0)       # :175-176 pass no epsilon

    @classmethod
    def reduce_metrics(cls, logging_outputs) -> None:
        """:210-261 (audio twin :206-253)."""
        nsentences = sum(log.get("nsentences", 1) for log in logging_outputs)
        sample_size = sum(log.get("sample_size", 1) for log in logging_outputs)
        for key in ("loss",) + tuple(cls.loss_keys) + ("logit_scale_exp",):
            total = sum(log.get(key, 0) for log in logging_outputs)
            metrics.log_scalar(key, total / sample_size, sample_size, round=3)
        metrics.log_scalar("nsentences", nsentences, 1, round=3)
        metrics.log_scalar("sample_size", sample_size, 1, round=3)
        for key in cls.acc_keys:
            if len(logging_outputs) > 0 and key in logging_outputs[0]:
                ncorrect = sum(log.get(key, 0) for log in logging_outputs)
                metrics.log_scalar(key.replace("ncorrect", "accuracy"), 100.0 * ncorrect / nsentences, nsentences, round=1)

    @staticmethod
    def logging_outputs_can_be_summed() -> bool:
        return True


@register_criterion("image_text_pretrain_loss")
class ImageTextPretrainLossCriterion(_PretrainCriterionBase):
    loss_keys = ("itc_loss", "dcl_text_loss", "dcl_image_loss", "dcl_vl_text_loss", "dcl_vl_image_loss")
    acc_keys = ("i2t_ncorrect", "t2i_ncorrect")

    def __init__(self, task, dcl_text_alpha=0.5, dcl_image_alpha=1.0, dcl_vl_text_alpha=0.5, dcl_vl_image_alpha=0.5,
                 dcl_logit_scale=2.5, label_smoothing=0.0):
        super().__init__(task)
        self.dcl_text_alpha = dcl_text_alpha
        self.dcl_image_alpha = dcl_image_alpha
        self.dcl_vl_text_alpha = dcl_vl_text_alpha
        self.dcl_vl_image_alpha = dcl_vl_image_alpha
        self.dcl_logit_scale = dcl_logit_scale
        self.label_smoothing = label_smoothing

    def forward(self, model, sample, reduce=True):
        """(loss, sample_size=1, logging_output) — image_text_pretrain_loss.py:76-162."""
        ni = sample["net_input"]
        src_tokens, src_images = ni["src_tokens"], ni["src_images"]
        text_logits, teacher_text = model(src_tokens=src_tokens, encoder_type="text")
        image_logits, teacher_image = model(src_images=src_images, encoder_type="image")
        text_all = gather_without_grad(text_logits) if dist.is_initialized() else text_logits.data
        image_all = gather_without_grad(image_logits) if dist.is_initialized() else image_logits.data
        with torch.no_grad():
            teacher_vl_text, teacher_vl_image = model(src_tokens=src_tokens, src_images=src_images, encoder_type="vl")
        student_text, _, _ = model(src_tokens=src_tokens, text_preserve_ids=ni["text_preserve_ids"], encoder_type="text")
        _, student_image, _ = model(src_images=src_images, image_preserve_ids=ni["image_preserve_ids"], encoder_type="image")
        student_vl_text, student_vl_image, _ = model(src_tokens=src_tokens, text_preserve_ids=ni["vl_text_preserve_ids"],
                                                     src_images=src_images, image_preserve_ids=ni["vl_image_preserve_ids"],
                                                     encoder_type="vl")
        logit_scale_exp = model(return_logit_scale=True)
        padding_masks = src_tokens.eq(1)
        dcl_text = self.compute_dcl_loss(student_text, teacher_text, ni["text_mask_indices"], padding_masks)
        dcl_image = self.compute_dcl_loss(student_image, teacher_image, ni["image_mask_indices"])
        dcl_vl_text = self.compute_dcl_loss(student_vl_text, teacher_vl_text, ni["vl_text_mask_indices"], padding_masks)
        dcl_vl_image = self.compute_dcl_loss(student_vl_image, teacher_vl_image, ni["vl_image_mask_indices"])
        itc, i2t_ok, t2i_ok = self.compute_itc_loss(image_logits, text_logits, image_all, text_all, logit_scale_exp)
        loss = itc + self.dcl_text_alpha * dcl_text + self.dcl_image_alpha * dcl_image + \
            self.dcl_vl_text_alpha * dcl_vl_text + self.dcl_vl_image_alpha * dcl_vl_image
        logging_output = {"loss": loss.data, "itc_loss": itc.data, "dcl_text_loss": dcl_text.data,
                          "dcl_image_loss": dcl_image.data, "dcl_vl_text_loss": dcl_vl_text.data,
                          "dcl_vl_image_loss": dcl_vl_image.data, "nsentences": sample["nsentences"], "sample_size": 1,
                          "i2t_ncorrect": i2t_ok, "t2i_ncorrect": t2i_ok, "logit_scale_exp": logit_scale_exp.data}
        return loss, 1, logging_output

    def compute_itc_loss(self, image_logits, text_logits, image_logits_all, text_logits_all, logit_scale_exp):
        return self._contrastive(image_logits, text_logits, image_logits_all, text_logits_all, logit_scale_exp)
