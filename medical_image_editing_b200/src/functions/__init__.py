from .vq_function import VQFunction, vq_lookup  # noqa: F401
from .embed_loss import EmbeddingLoss, cross_loss, labels_from_onehot  # noqa: F401
from .onehot import OneHotEncoder  # noqa: F401
from .kmeans import kmeans, kmeans_nchw, initialize_embed  # noqa: F401
from .norm_relu import InstanceNormReLU, instance_norm_relu, fuse_norm_relu_pairs, fuse_vq_tail  # noqa: F401
