"""Boundary constants of the reference (config.py:1-21)."""
grid_size = 20

# BGR, nearest-key lookup in PenaltyCalculator.get_penalty_colour
penalty_colour_gradient = {
    1.0000: (0, 0, 255), 0.9166: (0, 60, 255), 0.8333: (0, 88, 255), 0.7500: (0, 109, 255),
    0.6666: (0, 128, 255), 0.5833: (8, 145, 255), 0.5000: (0, 163, 249), 0.4166: (0, 183, 232),
    0.3333: (0, 202, 208), 0.1666: (0, 221, 176), 0.0833: (0, 239, 129), 0.0000: (0, 255, 15),
}
