from .preprocess import (preprocess_volumes, resize_array, resize_shape, to_training_volume,  # noqa: F401
                         training_loader_volume, inference_loader_volume, resample_to_target)
