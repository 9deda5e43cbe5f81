"""Column-name contract of the window / group-by steps.

Mirrors the reference's ``slaf.core.tabular_schema`` (slaf/core/tabular_schema.py:9-71): same
class name, same attributes, same default instance, so code written against the reference's
schema objects works unchanged (a reference ``DataSchema`` instance is accepted as well: only
attribute access is used).
"""

from __future__ import annotations

from dataclasses import dataclass


@dataclass
class DataSchema:
    group_key: str
    item_key: str
    value_key: str
    group_key_out: str | None = None
    item_list_key: str = "item_list"
    value_list_key: str | None = None


SLAF_LANCE_COO_SCHEMA = DataSchema(
    group_key="cell_integer_id",
    item_key="gene_integer_id",
    value_key="value",
    group_key_out="cell_integer_id",
    item_list_key="gene_sequence",
    value_list_key="expr_sequence",
)
