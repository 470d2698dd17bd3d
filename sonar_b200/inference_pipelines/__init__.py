from .text import TextToEmbeddingModelPipeline, precision_context  # noqa: F401
