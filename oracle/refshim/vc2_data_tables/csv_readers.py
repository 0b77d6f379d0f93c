"""
Stand-in for ``vc2_data_tables.csv_readers`` (import-only support for the
reference's level-constraint tables).  TEST INFRASTRUCTURE ONLY.
"""

import csv
import io
import re


def open_utf8(filename):
    return io.open(filename, "r", encoding="utf-8", newline="")


def is_ditto(string):
    return bool(re.match(u'^\\s*["“”‟″〃]\\s*$', string))


def read_lookup_from_csv(csv_filename, index_enum_type, namedtuple_type, type_conversions=None):
    type_conversions = type_conversions or {}
    out = {}
    with open_utf8(csv_filename) as f:
        rows = [
            row
            for row in csv.reader(f)
            if any(cell.strip() for cell in row)
            and not row[0].strip().startswith("#")
            and not (row[0].strip() == "" and all(
                (not c.strip()) or c.strip().startswith("#") for c in row))
        ]
    header = [c.strip() for c in rows[0]]
    last = {}
    for row in rows[1:]:
        record = dict(zip(header, row))
        try:
            index = index_enum_type(int(record["index"]))
        except ValueError:
            continue
        values = {}
        for field in namedtuple_type._fields:
            cell = record[field]
            if is_ditto(cell):
                values[field] = last[field]
            else:
                values[field] = type_conversions.get(field, lambda x: x)(cell)
        last = values
        out[index] = namedtuple_type(**values)
    return out
