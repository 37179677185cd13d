"""Action enum of the Tiler-Slider path (reference: explainrl/environment/state.py:29-45)."""
import enum


class Move(enum.Enum):
    """Sliding directions; the integer values index the action byte of the CUDA kernels."""
    UP = 0
    DOWN = 1
    LEFT = 2
    RIGHT = 3

    @classmethod
    def from_char(cls, direction: str):
        """'U','D','L','R' (case-insensitive) -> Move; anything else -> None."""
        return {"U": cls.UP, "D": cls.DOWN, "L": cls.LEFT, "R": cls.RIGHT}.get(direction.upper())

    @classmethod
    def from_int(cls, value: int):
        return cls(value)
