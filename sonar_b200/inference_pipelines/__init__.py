from .speech import (  # noqa: F401
    AudioToFbankDataPipelineBuilder,
    SpeechInferenceParams,
    SpeechToEmbeddingModelPipeline,
    SpeechToEmbeddingPipeline,
    SpeechToTextModelPipeline,
    SpeechToTextPipeline,
    read_tsv_audio_paths,
)  # noqa: F401
from .text import (  # noqa: F401
    EmbeddingToTextModelPipeline,
    TextToEmbeddingModelPipeline,
    TextToTextModelPipeline,
    precision_context,
)
