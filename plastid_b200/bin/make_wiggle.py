"""``make_wiggle``: export the count vectors of an alignment file under a mapping rule as browser
tracks, one file per strand (plastid/bin/make_wiggle.py:99-209).  The run-length / non-zero compaction
runs on the device (``pb_export_runs``); file contents equal the reference's."""
import argparse
import sys

from . import _cli


def get_rgb255(color):
    """'#RRGGBB' -> (r, g, b) (plastid/util/services/colors.py get_rgb255, hex strings only)."""
    c = color.lstrip("#")
    if len(c) != 6:
        raise ValueError("color must be an RGB hex string like '#0000FF'")
    return tuple(int(c[i:i + 2], 16) for i in (0, 2, 4))


def write_tracks(ga, outbase, track_name=None, color=None, output_format="bedgraph", window_size=100000):
    """make_wiggle.py:172-209: ``<outbase>_fw.wig`` ('+') and ``<outbase>_rc.wig`` ('-')."""
    name = outbase if track_name is None else track_name
    rgb = "%s,%s,%s" % get_rgb255(color) if color is not None else "0,0,0"
    outfn = ga.to_bedgraph if output_format == "bedgraph" else ga.to_variable_step
    paths = []
    for suffix, strand in (("fw", "+"), ("rc", "-")):
        path = "%s_%s.wig" % (outbase, suffix)
        with open(path, "w") as fh:
            outfn(fh, "%s_%s" % (name, suffix), strand, window_size=window_size, color=rgb)
        paths.append(path)
    return paths


def main(argv=sys.argv[1:]):
    """``make_wiggle -o OUTBASE --count_files ... [--output_format bedgraph|variable_step] [--normalize]`` with the
    reference's flags (plastid/bin/make_wiggle.py:99-209)."""
    parser = argparse.ArgumentParser(description=__doc__)
    _cli.add_base_args(parser)
    _cli.add_alignment_args(parser)
    parser.add_argument("-o", "--out", dest="outbase", required=True, metavar="FILENAME", help="Base name for output files")
    parser.add_argument("--window_size", default=100000, type=int, metavar="N")
    parser.add_argument("--color", default=None, help="RGB hex string '#NNNNNN'")
    parser.add_argument("-t", "--track_name", dest="track_name", default=None)
    parser.add_argument("--output_format", choices=("bedgraph", "variable_step"), default="bedgraph")
    args = parser.parse_args(argv)
    ga = _cli.genome_array_from_args(args)
    write_tracks(ga, args.outbase, args.track_name, args.color, args.output_format, args.window_size)
    _cli.finish_distributed()


if __name__ == "__main__":
    main()
