"""Generates tests/golden/fixtures.json.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

The reference (Rust) cannot be executed here, so the expected outputs are the
vectors the reference's own tests assert and the result tables its README
prints, transcribed below with their file:line.  The input tables are read
from the reference's data/*.csv.
"""
import csv
import json
import os

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixtures.json")


def table(path):
    with open(os.path.join(REF, path)) as f:
        rows = list(csv.reader(f))
    return {"names": rows[0], "rows": rows[1:]}


fixtures = {
    "tables": {
        "t1": table("data/test_data.csv"),
        "employee": table("data/employee.csv"),
        "rank": table("data/rank.csv"),
        "department": table("data/department.csv"),
    },
    "cases": {
        # src/physical_plan/projection.rs:88-121  (id + 1, name)
        "test_projection": {
            "cite": "src/physical_plan/projection.rs:88-121",
            "id_plus_1": [2, 3, 5, 6, 7, 8, 9, 10],
            "name": ["veeupup", "alex", "lynne", "alice", "bob", "jack", "cock", "primer"],
        },
        # src/physical_plan/selection.rs:126-178  ((id + 1) > 5 over id,name,age)
        "test_selection": {
            "cite": "src/physical_plan/selection.rs:126-178",
            "id": [5, 6, 7, 8, 9],
            "name": ["alice", "bob", "jack", "cock", "primer"],
        },
        # src/sql/planner.rs:664-680  select id,name,age from t1 where id > 1
        "sql_where_id_gt_1": {
            "cite": "src/sql/planner.rs:664-680",
            "id": [2, 4, 5, 6, 7, 8, 9],
            "name": ["alex", "lynne", "alice", "bob", "jack", "cock", "primer"],
            "age": [20, 18, 19, 20, 21, 22, 23],
        },
        # src/physical_plan/expression/unary.rs:123-170
        "test_abs_expression": {
            "cite": "src/physical_plan/expression/unary.rs:123-145",
            "score": [60.0, 90.1, 99.99, 81.1, 82.2, 83.3, 84.4, 85.5],
        },
        "test_sin_expression": {
            "cite": "src/physical_plan/expression/unary.rs:147-170",
            "score": [-0.3048106211022167, 0.8447976840197418, -0.5149633680424761,
                      -0.5492019627147913, 0.49565689358989423, 0.9988580516952367,
                      0.4104993826174394, -0.6264561960895026],
        },
        # src/physical_plan/limit.rs:67-90, offset.rs:69-92
        "test_physical_offset": {"cite": "src/physical_plan/offset.rs:69-92", "id": [7, 8, 9]},
        # BASELINE.json config 1 (derived)
        "config1": {
            "cite": "BASELINE.json configs[0]; selection.rs:58-107, projection.rs:43-70",
            "sql": "select id, age+100 from t1 where id < 9",
            "names": ["id", "age + 100"],
            "rows": [[1, 123], [2, 120], [4, 118], [5, 119], [6, 120], [7, 121], [8, 122]],
        },
        # README.md:70-76
        "readme_limit_offset": {
            "cite": "README.md:70-76",
            "sql": "select id, name, age + 100 from t1 where id < 9 limit 3 offset 2",
            "rows": [[4, "lynne", 118], [5, "alice", 119], [6, "bob", 120]],
        },
        # README.md:77-85 three-way join, rows in printed order
        "readme_join": {
            "cite": "README.md:77-85",
            "sql": "select id, name, rank_name, department_name from employee "
                   "join rank on employee.rank = rank.id "
                   "join department on employee.department_id = department.id",
            "rows": [[2, "lynne", "master", "IT"], [1, "vee", "diamond", "IT"],
                     [3, "Alex", "master", "Marketing"], [4, "jack", "diamond", "Marketing"],
                     [5, "mike", "grandmaster", "Human Resource"]],
        },
        # README.md:105-111 group by id % 3 (row order is HashMap-random)
        "readme_groupby": {
            "cite": "README.md:105-111",
            "sql": "select count(id), sum(age), sum(score), avg(score), max(score), min(score) "
                   "from t1 group by id % 3",
            "names": ["count(id)", "sum(age)", "sum(score)", "avg(score)", "max(score)", "min(score)"],
            "rows": [[3, 61.0, 255.6, 85.2, 90.1, 81.1],
                     [3, 62.0, 243.29000000000002, 81.09666666666668, 99.99, 60.0],
                     [2, 43.0, 167.7, 83.85, 85.5, 82.2]],
        },
    },
}

with open(OUT, "w") as f:
    json.dump(fixtures, f, indent=1)
print("wrote", OUT)
