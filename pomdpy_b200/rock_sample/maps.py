"""RockSample map registry and the ASCII map format of the reference.

Format (reference pomdpy/util/config_parser.py:14-21, examples/rock_sample/rock_model.py:110-138): first line
"<rows> <cols>", then `rows` lines of `cols` characters: '.' empty, 'o' rock (numbered in row-major scan
order), 'G' goal, 'S' start (an empty cell), 'X' obstacle.  The shipped maps are stored here as geometry
(grid size, start, rock coordinates; the goal is the last column) and rendered on demand."""
import os

# name -> (rows, cols, start (i, j), rock coordinates in scan order)
GEOMETRY = {
    "map-7-2": (7, 7, (3, 0), [(2, 3), (6, 1)]),
    "map-7-3": (7, 7, (3, 0), [(1, 4), (4, 2), (6, 1)]),
    "map-7-4": (7, 7, (3, 0), [(0, 0), (1, 4), (5, 5), (6, 1)]),
    "map-7-8": (7, 7, (3, 0), [(0, 1), (0, 4), (1, 0), (3, 1), (3, 4), (4, 1), (5, 0), (6, 3)]),
    "map-7-8-7": (7, 8, (3, 0), [(0, 1), (1, 5), (2, 2), (2, 3), (3, 6), (5, 0), (5, 3), (6, 2)]),
    "map-10-10": (10, 10, (3, 0), [(0, 1), (1, 4), (1, 7), (4, 5), (4, 7), (5, 1), (6, 2), (6, 7), (8, 5), (9, 1)]),
    "map-11-11": (11, 11, (3, 0), [(0, 1), (1, 4), (1, 7), (4, 5), (4, 7), (5, 1), (6, 2), (6, 7), (8, 5), (9, 1),
                                   (10, 7)]),
    "map-12-12": (12, 12, (3, 0), [(0, 1), (0, 7), (1, 4), (1, 7), (2, 6), (4, 5), (4, 9), (6, 2), (6, 9), (8, 7),
                                   (9, 1), (11, 6)]),
    "map-15-15": (15, 15, (3, 0), [(0, 1), (0, 10), (1, 4), (2, 6), (4, 5), (4, 12), (6, 12), (7, 3), (8, 7), (9, 1),
                                   (11, 6), (12, 3), (12, 12), (14, 2), (14, 10)]),
}
DEFAULT_MAP = "map-7-8"            # the reference's rock_sample_config.json default


def _canonical(name):
    name = os.path.basename(str(name))
    return name[:-4] if name.endswith(".txt") else name


def render(name):
    """The ASCII text of a registered map, in the reference's file format."""
    rows, cols, start, rocks = GEOMETRY[_canonical(name)]
    grid = [["."] * (cols - 1) + ["G"] for _ in range(rows)]
    grid[start[0]][start[1]] = "S"
    for (i, j) in rocks:
        grid[i][j] = "o"
    return "%d %d\n" % (rows, cols) + "\n".join("".join(r) for r in grid) + "\n"


def parse(text):
    """ASCII map -> (rows, cols, list of row strings)."""
    lines = text.splitlines()
    dims = lines[0].split()
    rows, cols = int(dims[0]), int(dims[1])
    body = [ln.strip() for ln in lines[1:] if ln.strip()]
    if len(body) != rows or any(len(ln) != cols for ln in body):
        raise ValueError("map body does not match its %dx%d header" % (rows, cols))
    bad = set("".join(body)) - set(".oGSX")
    if bad:
        raise ValueError("unknown map characters: %r" % sorted(bad))
    return rows, cols, body


def load(name_or_path):
    """A registered map name ("map-15-15", "map-15-15.txt") or a path to a file in the same format."""
    if _canonical(name_or_path) in GEOMETRY:
        return parse(render(name_or_path))
    with open(name_or_path, "r") as f:
        return parse(f.read())
