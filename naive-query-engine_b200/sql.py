"""The SQL surface the reference's planner accepts, restated so that `NaiveDB.run_sql()` can drive the GPU operators
(SURVEY.md 8f-1).  This is glue in front of the hot path, not part of it: the reference parses with sqlparser 0.9
(GenericDialect, src/sql/parser.rs:19-25) and plans with SQLPlanner (src/sql/planner.rs:45-380); neither is GPU work.

Only what SQLPlanner handles is parsed; everything it answers with `unimplemented!()` / `todo!()` / `Err(..)` is
answered the same way here (NqeError kinds Panic / NotImplemented / PlanError):

    SELECT item [, item]*  FROM t [, t]*  { [INNER] JOIN t ON expr | LEFT|RIGHT [OUTER] JOIN t ON expr | CROSS JOIN t }*
    [WHERE expr] [GROUP BY expr [, expr]*] [ORDER BY ... (ignored, planner.rs:159-162)] [LIMIT n] [OFFSET n [ROW|ROWS]]

    item := * | expr            (an alias is `unimplemented!()`, planner.rs:144)
    expr := literal | ident | ident.ident | expr op expr | @ expr | CAST(expr AS type) | name(expr)
            op in  = != <> < <= > >= + - * / % AND OR;  a parenthesised expression is `todo!()` (Expr::Nested, :474)

Logical expressions are tuples:
    ("col", table|None, name)  ("lit", kind, value)  ("bin", Operator, l, r)  ("un", "Abs", arg)
    ("cast", expr, arrow_type)  ("agg", "count|sum|avg|min|max", arg)  ("wildcard",)
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import List, Optional

from ._ffi import NqeError

_TOKEN = re.compile(r"""\s*(?:
    (?P<num>\d+\.\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?|\d+(?:[eE][+-]?\d+)?)
  | (?P<str>'(?:[^']|'')*')
  | (?P<qid>"[^"]*"|`[^`]*`)
  | (?P<id>[A-Za-z_][A-Za-z_0-9]*)
  | (?P<op><>|!=|<=|>=|[=<>+\-*/%(),.;@])
)""", re.X)

_BINOPS = {"=": "Eq", "!=": "NotEq", "<>": "NotEq", "<": "Lt", "<=": "LtEq", ">": "Gt", ">=": "GtEq", "+": "Plus",
           "-": "Minus", "*": "Multiply", "/": "Divide", "%": "Modulos", "AND": "And", "OR": "Or"}
# sqlparser precedences: OR 5, AND 10, comparisons 20, + - 30, * / % 40
_PREC = {"Or": 5, "And": 10, "Eq": 20, "NotEq": 20, "Lt": 20, "LtEq": 20, "Gt": 20, "GtEq": 20, "Plus": 30, "Minus": 30,
         "Multiply": 40, "Divide": 40, "Modulos": 40}
_KEYWORDS = {"SELECT", "FROM", "WHERE", "GROUP", "BY", "ORDER", "LIMIT", "OFFSET", "JOIN", "INNER", "LEFT", "RIGHT", "OUTER",
             "CROSS", "ON", "AND", "OR", "AS", "CAST", "NULL", "TRUE", "FALSE", "ROW", "ROWS", "ASC", "DESC", "FULL", "NATURAL", "USING"}


@dataclass
class Select:
    projection: list
    from_: list                      # [(table name, [(join_type, table name, on expr | None)])]
    selection: Optional[tuple] = None
    group_by: list = field(default_factory=list)
    limit: Optional[tuple] = None
    offset: Optional[tuple] = None


def _tokens(sql: str):
    out, pos = [], 0
    sql = sql.strip()
    while pos < len(sql):
        m = _TOKEN.match(sql, pos)
        if not m or m.end() == pos:
            raise NqeError(8, f"sql parser error: unexpected character at {sql[pos:pos + 10]!r}")  # ParserError
        pos = m.end()
        kind = m.lastgroup
        text = m.group(kind)
        if kind == "id" and text.upper() in _KEYWORDS:
            out.append(("kw", text.upper()))
        elif kind == "id":
            out.append(("id", text.lower()))          # normalize_ident: unquoted identifiers are lower-cased
        elif kind == "qid":
            out.append(("id", text[1:-1]))
        elif kind == "str":
            out.append(("str", text[1:-1].replace("''", "'")))
        else:
            out.append((kind, text))
    return out


class _Parser:
    def __init__(self, sql: str):
        self.t = _tokens(sql)
        self.i = 0

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else ("eof", "")

    def take(self):
        tok = self.peek()
        self.i += 1
        return tok

    def kw(self, *words) -> bool:
        for k, w in enumerate(words):
            if self.peek(k) != ("kw", w):
                return False
        self.i += len(words)
        return True

    def expect(self, kind, text=None):
        tok = self.take()
        if tok[0] != kind or (text is not None and tok[1] != text):
            raise NqeError(8, f"sql parser error: expected {text or kind}, found {tok[1]!r}")
        return tok[1]

    # ---- expressions (precedence climbing)
    def expr(self, min_prec=0):
        left = self.prefix()
        while True:
            tok = self.peek()
            name = _BINOPS.get(tok[1]) if tok[0] in ("op", "kw") else None
            if name is None or _PREC[name] <= min_prec:
                return left
            self.take()
            right = self.expr(_PREC[name])
            left = ("bin", name, left, right)

    def prefix(self):
        kind, text = self.take()
        if kind == "num":
            try:
                return ("lit", "Int64", int(text))          # planner.rs:451-454: i64 first, else f64
            except ValueError:
                return ("lit", "Float64", float(text))
        if kind == "str":
            return ("lit", "Utf8", text)
        if kind == "kw" and text in ("TRUE", "FALSE"):
            return ("lit", "Boolean", text == "TRUE")
        if kind == "kw" and text == "NULL":
            return ("lit", "Null", None)
        if kind == "op" and text == "@":                     # UnaryOperator::PGAbs -> Abs (planner.rs:556-565)
            return ("un", "Abs", self.expr(50))
        if kind == "op" and text in ("-", "+"):
            self.expr(50)
            raise NqeError(5, "not implemented: unary operator")       # unimplemented!() planner.rs:559
        if kind == "op" and text == "(":
            raise NqeError(5, "not yet implemented: nested expression")  # Expr::Nested -> todo!() planner.rs:524
        if kind == "kw" and text == "CAST":
            self.expect("op", "(")
            e = self.expr()
            self.expect("kw", "AS")
            ty = self.take()[1].upper()
            if self.peek() == ("op", "("):                    # VARCHAR(20), FLOAT(8), ...
                while self.take() != ("op", ")"):
                    pass
            self.expect("op", ")")
            return ("cast", e, ty)
        if kind == "id":
            if self.peek() == ("op", "("):                    # function call
                self.take()
                args = []
                if self.peek() != ("op", ")"):
                    if self.peek() == ("op", "*"):
                        self.take()
                        args.append(("wildcard",))
                    else:
                        args.append(self.expr())
                    while self.peek() == ("op", ","):
                        self.take()
                        args.append(self.expr())
                self.expect("op", ")")
                return ("call", text, args)
            if self.peek() == ("op", ".") and self.peek(1)[0] == "id":
                self.take()
                name = self.take()[1]
                if self.peek() == ("op", "."):
                    raise NqeError(4, "compound identifier with more than two parts")  # Err(NotImplemented) planner.rs:472
                return ("col", text, name)
            return ("col", None, text)
        raise NqeError(8, f"sql parser error: unexpected {text!r}")

    # ---- statement
    def table(self) -> str:
        name = self.expect("id")
        while self.peek() == ("op", "."):
            self.take()
            name += "." + self.expect("id")
        if self.peek()[0] == "id" or self.kw("AS"):          # alias: TableFactor::Table {alias} is ignored (`name, ..`)
            if self.peek()[0] == "id":
                self.take()
        return name

    def select(self) -> Select:
        if not self.kw("SELECT"):
            raise NqeError(5, "not implemented: only SELECT statements")  # statement_to_plan: `_ => unimplemented!()`
        items = []
        while True:
            if self.peek() == ("op", "*"):
                self.take()
                items.append(("wildcard",))
            else:
                e = self.expr()
                if self.kw("AS") or self.peek()[0] == "id":
                    raise NqeError(5, "not implemented: select item with an alias")  # planner.rs:144
                items.append(e)
            if self.peek() != ("op", ","):
                break
            self.take()
        if not self.kw("FROM"):
            raise NqeError(5, "not yet implemented: support select with no from")      # planner.rs:196
        from_ = []
        while True:
            rel = self.table()
            joins = []
            while True:
                if self.kw("CROSS", "JOIN"):
                    joins.append(("Cross", self.table(), None))
                    continue
                jt = None
                if self.kw("INNER", "JOIN") or self.kw("JOIN"):
                    jt = "Inner"
                elif self.kw("LEFT", "OUTER", "JOIN") or self.kw("LEFT", "JOIN"):
                    jt = "Left"
                elif self.kw("RIGHT", "OUTER", "JOIN") or self.kw("RIGHT", "JOIN"):
                    jt = "Right"
                elif self.kw("FULL", "OUTER", "JOIN") or self.kw("FULL", "JOIN") or self.kw("NATURAL"):
                    raise NqeError(4, "join operator")                                 # `_other => Err(NotImplemented)` :234
                if jt is None:
                    break
                t = self.table()
                if self.kw("ON"):
                    joins.append((jt, t, self.expr()))
                elif self.kw("USING"):
                    raise NqeError(4, "join constraint")                               # `_ => Err(NotImplemented)` :281
                else:
                    joins.append((jt, t, None))
            from_.append((rel, joins))
            if self.peek() != ("op", ","):
                break
            self.take()
        sel = Select(items, from_)
        if self.kw("WHERE"):
            sel.selection = self.expr()
        if self.kw("GROUP", "BY"):
            sel.group_by.append(self.expr())
            while self.peek() == ("op", ","):
                self.take()
                sel.group_by.append(self.expr())
        if self.kw("ORDER", "BY"):                            # parsed and ignored (planner.rs:159-162)
            while True:
                self.expr()
                if not (self.kw("ASC") or self.kw("DESC")):
                    pass
                if self.peek() != ("op", ","):
                    break
                self.take()
        # sqlparser accepts LIMIT and OFFSET in either order
        for _ in range(2):
            if self.kw("LIMIT"):
                sel.limit = self.expr()
            elif self.kw("OFFSET"):
                sel.offset = self.expr()
                _ = self.kw("ROWS") or self.kw("ROW")
        if self.peek() == ("op", ";"):
            self.take()
        if self.peek()[0] != "eof":
            raise NqeError(8, f"sql parser error: unexpected {self.peek()[1]!r} after the statement")
        return sel


def parse(sql: str) -> Select:
    """SQLParser::parse (src/sql/parser.rs:19-25): one statement."""
    return _Parser(sql).select()
