"""capdec_b200 — B200-native (sm_100a) implementation of the CapDec training-step hot path.

Public surface mirrors the reference (train.py / gpt2_prefix.py): ClipCaptionModel, ClipCaptionPrefix, MappingType,
MLP, TransformerMapper, noise_injection; plus Trainer (the fused, CUDA-graph'ed train step), fit.train (the reference's train() loop on it) and AdamW /
get_linear_schedule_with_warmup with the reference's HuggingFace semantics.  There is no CPU fallback: the CUDA
library (capdec_b200/libcapdec_b200.so, built by `python -m capdec_b200.build`) is required.
"""
from .model import (ClipCaptionModel, ClipCaptionPrefix, GPT2Config, GPT2LMHead, MappingType, MLP, TransformerEncoderDecoder, TransformerMapper,
                    noise_injection)
from .optim import AdamW, get_linear_schedule_with_warmup
from .trainer import Trainer
from .data import DeviceCaptionDataset
from .decode import BeamDecoder, generate2, generate_beam, generate_beam_batch, generate_beam_ids, generate_greedy_ids
from .bridger import ModalityBridger, get_map_to_text_space_using_modality_bridger
from . import ops
from . import fit

__all__ = ["ModalityBridger", "get_map_to_text_space_using_modality_bridger", "ClipCaptionModel", "ClipCaptionPrefix", "GPT2Config", "GPT2LMHead", "MappingType", "MLP",
           "TransformerMapper", "TransformerEncoderDecoder", "noise_injection", "AdamW", "get_linear_schedule_with_warmup", "Trainer", "ops",
           "DeviceCaptionDataset", "BeamDecoder", "generate_beam", "generate_beam_batch", "generate_beam_ids", "generate2", "generate_greedy_ids", "fit"]
