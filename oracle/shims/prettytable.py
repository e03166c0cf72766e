"""Stand-in for ``prettytable.PrettyTable`` (lib/dataset/h36m.py:4): collects rows and prints them plainly."""


class PrettyTable:
    def __init__(self, field_names=None, **kwargs):
        self.field_names = list(field_names or [])
        self.rows = []

    def add_row(self, row, **kwargs):
        self.rows.append(list(row))

    def add_column(self, name, column, **kwargs):
        self.field_names.append(name)
        for i, v in enumerate(column):
            if i >= len(self.rows):
                self.rows.append([])
            self.rows[i].append(v)

    def __str__(self):
        lines = [" | ".join(str(f) for f in self.field_names)]
        lines += [" | ".join(str(c) for c in r) for r in self.rows]
        return "\n".join(lines)
