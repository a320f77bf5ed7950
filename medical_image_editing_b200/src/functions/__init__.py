from .vq_function import VQFunction, vq_lookup  # noqa: F401
from .embed_loss import EmbeddingLoss, cross_loss, labels_from_onehot  # noqa: F401
from .onehot import OneHotEncoder  # noqa: F401
from .kmeans import kmeans, kmeans_nchw, initialize_embed  # noqa: F401
