from .parallel_state import (
    all_reduce_min,
    all_reduce_tp,
    divide,
    get_tp_comm,
    get_tp_rank,
    get_tp_world_size,
    init_tensor_parallel,
    destroy_tensor_parallel,
)

__all__ = ["all_reduce_min", "all_reduce_tp", "divide", "get_tp_comm", "get_tp_rank", "get_tp_world_size",
           "init_tensor_parallel", "destroy_tensor_parallel"]
