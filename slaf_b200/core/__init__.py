from .tabular_schema import SLAF_LANCE_COO_SCHEMA, DataSchema  # noqa: F401
