"""Error types of the interface mirror (kept free of torch / CUDA imports)."""


class DimensionError(Exception):
    """NLPModels.DimensionError(name, dim_expected, dim_found)."""

    def __init__(self, name, dim_expected, dim_found):
        super().__init__(f"DimensionError: Input {name} should have length {dim_expected} not {dim_found}")
        self.name, self.dim_expected, self.dim_found = name, dim_expected, dim_found
