"""``QuantitySpec`` / ``ChunkSpec`` -- same fields as ``mlmc/quantity/quantity_spec.py:7-28``."""
from dataclasses import dataclass
from typing import List, Optional, Tuple, Union

import numpy as np


@dataclass(eq=False)
class QuantitySpec:
    name: str
    unit: str
    shape: Tuple[int, int]
    times: List[float]
    locations: Union[List[str], List[Tuple[float, float, float]]]

    def __eq__(self, other):
        return (self.name, self.unit) == (other.name, other.unit) \
            and np.array_equal(self.shape, other.shape) \
            and np.array_equal(self.times, other.times) \
            and not (set(self.locations) - set(other.locations))

    __hash__ = object.__hash__


@dataclass
class ChunkSpec:
    chunk_id: Optional[int] = None
    chunk_slice: Optional[slice] = None
    level_id: Optional[int] = None
