"""naive-query-engine_b200 -- B200-native (sm_100a) execution layer for
naive-query-engine's physical_plan pipeline.

    csrc/              CUDA kernels + the C ABI (include/nqe.h) -> libnqe_b200.so
    _ffi.py            ctypes binding of that ABI
    device.py          Arrow RecordBatch <-> HBM tables
    physical_plan.py   host mirror of the reference's PhysicalPlan / PhysicalExpr /
                       AggregateOperator interface
    sql.py, db.py      `NaiveDB.run_sql()`: the reference's SQL surface and planner wiring in front of those nodes
    distributed.py     multi-GPU plans (torch.distributed)
    synth.py           deterministic synthetic benchmark tables

The directory name contains a hyphen; import it with
`importlib.import_module("naive-query-engine_b200")` or via the `nqe_b200`
alias module at the repository root.
"""
from ._ffi import LIB_PATH, NqeError, load  # noqa: F401
from .device import Context, DeviceTable, MultiContext  # noqa: F401
from .physical_plan import (Avg, ColumnExpr, Count, CsvTable, HashJoin, Max, MemTable, Min,  # noqa: F401
                            PhysicalAggregatePlan, PhysicalBinaryExpr, PhysicalCastExpr, PhysicalLimitPlan,
                            PhysicalExpr, PhysicalLiteralExpr, PhysicalOffsetPlan, PhysicalPlan, PhysicalUnaryExpr,
                            ProjectionPlan, ScalarValue, ScanPlan, SelectionPlan, Sum)
from .db import CsvConfig, NaiveDB, QueryPlanner, SQLPlanner  # noqa: F401,E402
