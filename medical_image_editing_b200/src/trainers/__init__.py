from .ddp import DataParallelVQTrainer, broadcast_module_state, all_reduce_gradients  # noqa: F401
