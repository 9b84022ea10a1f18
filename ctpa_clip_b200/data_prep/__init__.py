from .preprocess import preprocess_volumes, resize_array, resize_shape, to_training_volume  # noqa: F401
