"""Boundary constants of the reference interface (reference config.py:1-21): the cell size every stage shares and
the penalty colour look-up used by PenaltyCalculator.get_penalty_colour (nearest key, colours are BGR)."""
grid_size = 20

# threshold keys in descending order and the green / red components of their colours (blue is 0 except one entry)
_KEYS = (1.0, 0.9166, 0.8333, 0.75, 0.6666, 0.5833, 0.5, 0.4166, 0.3333, 0.1666, 0.0833, 0.0)
_GREEN = (0, 60, 88, 109, 128, 145, 163, 183, 202, 221, 239, 255)
_RED = (255, 255, 255, 255, 255, 255, 249, 232, 208, 176, 129, 15)
_BLUE = (0, 0, 0, 0, 0, 8, 0, 0, 0, 0, 0, 0)
penalty_colour_gradient = {k: (b, g, r) for k, b, g, r in zip(_KEYS, _BLUE, _GREEN, _RED)}
