"""Mapping-rule plugin descriptors for plastid's entry-point registry
(``plastid/util/scriptlib/argparsers.py:505-534``, ``docs/source/devinfo/entrypoints.rst``): each
``plastid.mapping_rules`` entry point resolves to a dict with ``name`` / ``bamfunc`` / ``help``; the
script plumbing calls ``functools.partial(bamfunc, args=args)()`` (``argparsers.py:695-696``)."""
from .map_factories import (CenterMapFactory, FivePrimeMapFactory, ThreePrimeMapFactory,
                            VariableFivePrimeMapFactory)

fiveprime = dict(name="b200_fiveprime", help="Map reads at --offset from their 5' end (B200 kernels)",
                 bamfunc=lambda args=None: FivePrimeMapFactory(int(getattr(args, "offset", 0))))
threeprime = dict(name="b200_threeprime", help="Map reads at --offset from their 3' end (B200 kernels)",
                  bamfunc=lambda args=None: ThreePrimeMapFactory(int(getattr(args, "offset", 0))))
center = dict(name="b200_center", help="Trim --nibble from both ends, spread the read over the rest (B200 kernels)",
              bamfunc=lambda args=None: CenterMapFactory(int(getattr(args, "nibble", 0))))
fiveprime_variable = dict(name="b200_fiveprime_variable",
                          help="Per-read-length 5' offsets from the file given as --offset (B200 kernels)",
                          bamfunc=lambda args=None: VariableFivePrimeMapFactory.from_file(str(args.offset)))
device_option = dict(name="device", default="cuda", help="CUDA device the B200 mapping rules run on")
