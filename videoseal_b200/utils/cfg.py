"""`videoseal.utils.cfg` of the reference, same names: the implementation lives in videoseal_b200/cfg.py."""
from ..cfg import (get_config_from_checkpoint, maybe_download_checkpoint, resolve_config_path, setup_model,  # noqa: F401
                   setup_model_from_checkpoint, setup_model_from_model_card)
