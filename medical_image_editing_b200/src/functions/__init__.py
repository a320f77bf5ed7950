from .vq_function import VQFunction, vq_lookup  # noqa: F401
from .embed_loss import EmbeddingLoss, cross_loss, labels_from_onehot  # noqa: F401
