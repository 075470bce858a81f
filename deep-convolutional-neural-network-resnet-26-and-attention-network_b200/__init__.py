"""B200-native ResNet-26 + attention-MIL hot path (drop-in for the reference's `model.Attention`).

    import importlib
    mil = importlib.import_module("deep-convolutional-neural-network-resnet-26-and-attention-network_b200")
    classifier = mil.Attention(n_classes=3, class_weights=None).cuda()      # gbm/classify_combined.py:518
    output = classifier(bag, label); output['loss'].backward()              # gbm/classify_combined.py:432,447

The compute lives in libmil_b200.so (csrc/, C ABI in include/mil_b200.h); build it with `build()`.
"""
from ._build import build, LIB_PATH  # noqa: F401
from .distributed import BagGroup, SlideGroup  # noqa: F401
from .staging import BagStager  # noqa: F401
from .graph import GraphedStep  # noqa: F401
from .ingest import TileIngest  # noqa: F401
from .optim import FusedAdam, flatten_parameters, load_checkpoint, save_checkpoint, set_stage  # noqa: F401
from .export import export_attention_maps, minmax_normalize, top_tiles, write_dla  # noqa: F401
from .model import Attention, BasicResBlock, ContextLayer, CrossEntropyWithProbs, ResNet  # noqa: F401
from .wide import AltBasicBlock, AltResNet, WideAttention  # noqa: F401
from . import _lib, model, synth, wide  # noqa: F401

__all__ = ["Attention", "WideAttention", "AltResNet", "AltBasicBlock", "BagGroup", "SlideGroup", "BagStager", "GraphedStep", "TileIngest", "FusedAdam", "flatten_parameters", "set_stage", "save_checkpoint", "load_checkpoint", "ResNet", "BasicResBlock", "ContextLayer", "CrossEntropyWithProbs", "build",
           "LIB_PATH"]
