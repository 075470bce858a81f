"""Synthetic-bag generator: lives in the package (`<pkg>/synth.py`, numpy only) so that bench.py's GPU arm can
use it without importing anything from oracle/; re-exported here for the tests and the golden generator."""
import importlib

_m = importlib.import_module("deep-convolutional-neural-network-resnet-26-and-attention-network_b200.synth")
make_bag = _m.make_bag
make_drop_mask = _m.make_drop_mask
