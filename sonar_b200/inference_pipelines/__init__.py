from .speech import SpeechToEmbeddingModelPipeline, SpeechToTextModelPipeline  # noqa: F401
from .text import (  # noqa: F401
    EmbeddingToTextModelPipeline,
    TextToEmbeddingModelPipeline,
    TextToTextModelPipeline,
    precision_context,
)
