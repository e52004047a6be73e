"""`redux (-c | -d) [-i <input file>] [-o <output file>]` -- the reference CLI (src/main.rs:84-121) on the
B200 path: same flags, same fixed model AdaptiveTreeModel(Parameters(8, 30, 32)) (src/main.rs:108), same
stderr summary lines, same exit codes (1 usage, 2 open error, 3 codec error), byte-identical headerless files.

    python -m redux_b200.cli -c -i book1 -o book1.rdx
"""
import sys

import redux_b200 as rb

USAGE = "Usage: redux (-c | -d) [-i <input file>] [-o <output file>]"


def parse(argv):
    """Options::from_args (src/main.rs:36-61). Returns (compress, input, output) or None."""
    compress, inp, out = None, None, None
    it = iter(argv)
    for arg in it:
        if arg == "-c":
            compress = True
        elif arg == "-d":
            compress = False
        elif arg in ("-i", "-o"):
            val = next(it, None)
            if val is None:
                return None
            if arg == "-i":
                inp = val
            else:
                out = val
        else:
            return None
    return None if compress is None else (compress, inp, out)


def main(argv=None, stdin=None, stdout=None, stderr=None):
    argv = sys.argv[1:] if argv is None else argv
    stdin = stdin or sys.stdin.buffer
    stdout = stdout or sys.stdout.buffer
    stderr = stderr or sys.stderr
    opts = parse(argv)
    if opts is None:
        print(USAGE, file=stderr)
        return 1
    compress, inp, out = opts
    try:
        fin = stdin if inp is None else open(inp, "rb")
    except OSError as e:
        print("Error while opening input file %s: %s" % (inp, e), file=stderr)
        return 2
    try:
        fout = stdout if out is None else open(out, "wb")
    except OSError as e:
        print("Error while opening output file %s: %s" % (out, e), file=stderr)
        return 2
    model = rb.AdaptiveTreeModel.new(rb.Parameters.new(8, 30, 32))
    try:
        if compress:
            try:
                i, o = rb.compress(fin, fout, model)
            except rb.ReduxError as e:
                print("Compression error: %s" % e, file=stderr)
                return 3
            print("Compressed %d bytes into %d bytes, ratio: %.3f" % (i, o, i / o), file=stderr)
        else:
            try:
                i, o = rb.decompress(fin, fout, model)
            except rb.ReduxError as e:
                print("Decompression error: %s" % e, file=stderr)
                return 3
            print("Decompressed %d bytes from %d bytes, ratio: %.3f" % (o, i, o / i), file=stderr)
    finally:
        if out is not None:
            fout.close()
        if inp is not None:
            fin.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
