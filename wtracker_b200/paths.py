"""Locations of the committed model files."""
from pathlib import Path

REPO_ROOT = Path(__file__).resolve().parents[1]
MODELS_DIR = REPO_ROOT / "models"
RESMLP_100 = str(MODELS_DIR / "ResMLP(imaging-100ms_pred-40ms_moving-50ms).pt")
RESMLP_200 = str(MODELS_DIR / "ResMLP(imaging-200ms_pred-40ms_moving-50ms).pt")
