from wtracker_b200.neural.config import IOConfig
from wtracker_b200.neural.mlp import MLPLayer, MlpBlock, RMLP, WormPredictor, load_worm_predictor

__all__ = ["IOConfig", "MLPLayer", "MlpBlock", "RMLP", "WormPredictor", "load_worm_predictor"]
