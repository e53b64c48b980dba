"""asr_chinese_e2e_b200 -- B200-native CTC loss+grad hot path for zqs01/ASR_chinese_e2e.

Public surface: ``ctc_loss_b200`` / ``CTCLossB200`` (the op), ``JointCTCAttention`` (model-side
glue for ``TransformerOffical.cal_metrics``), ``sharded_ctc_loss`` (batch-sharded multi-GPU loss).
Importing the package does not load the CUDA library; the first call does, and raises if
libctcb200.so is missing -- there is no CPU fallback.
"""
from .ce import attention_ce_b200  # noqa: F401
from .ctc import CTCLossB200, ctc_greedy_cer_b200, ctc_loss_b200  # noqa: F401
from .ddp import DistributedWrapper, init_from_env, shard_batch  # noqa: F401
from .head import ctc_head_loss_b200  # noqa: F401
from .joint import JointCTCAttention, Pack  # noqa: F401
from .metrics import seq_cer_b200, seq_edit_distance_b200  # noqa: F401
from .masks import get_attn_key_pad_mask, get_attn_pad_mask, get_non_pad_mask, get_subsequent_mask  # noqa: F401
from .sharded import combine_equal_shards, combine_sharded_mean, sharded_ctc_loss  # noqa: F401

__all__ = ["CTCLossB200", "ctc_loss_b200", "ctc_head_loss_b200", "attention_ce_b200", "ctc_greedy_cer_b200", "JointCTCAttention", "Pack", "combine_equal_shards", "combine_sharded_mean",
           "sharded_ctc_loss", "DistributedWrapper", "init_from_env", "shard_batch", "get_non_pad_mask",
           "get_subsequent_mask", "get_attn_key_pad_mask", "get_attn_pad_mask", "seq_cer_b200", "seq_edit_distance_b200"]
