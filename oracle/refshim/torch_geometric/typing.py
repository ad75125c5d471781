from typing import Optional, Tuple, Union
import torch
Adj = Union[torch.Tensor]
Size = Optional[Tuple[int, int]]
OptTensor = Optional[torch.Tensor]
