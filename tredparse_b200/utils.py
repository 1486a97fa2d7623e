"""
Small host-side helpers mirroring the slice of ``tredparse/utils.py`` that the hot path's API uses:
``InputParams`` (utils.py:24-51), ``datafile`` (:77-78), ``mkdir`` (:81-102), ``listify`` (:105-106),
``DefaultHelpParser`` (:54-58).  S3 / shell helpers are out of scope.
"""
import argparse
import logging
import os
import os.path as op
import shutil
import sys

DATADIR = op.join(op.dirname(op.abspath(__file__)), "data")


def datafile(name):
    return op.join(DATADIR, name)


class InputParams:
    """All inputs of one (sample, locus) problem, as the reference's BamParser expects them."""
    KWARGS_LOG = "log"

    def __init__(self, bam, READLEN, repo, tredName, gender="Unknown", depth=30, clip=False,
                 alts=True, repeatpairs=False, **kwargs):
        self.bam = bam
        self.READLEN = READLEN
        self.tredName = tredName
        self.gender = gender
        self.depth = depth
        self.tred = repo.get(tredName)
        self.clip = clip
        self.alts = alts
        self.repeatpairs = repeatpairs
        self.kwargs = kwargs
        self.ref = repo.ref

    def getLogLevel(self, defaultLevel="INFO"):
        name = self.kwargs.get(InputParams.KWARGS_LOG, defaultLevel)
        return getattr(logging, str(name).upper(), defaultLevel)


class DefaultHelpParser(argparse.ArgumentParser):
    def error(self, message):
        sys.stderr.write("error: {}\n\n".format(message))
        sys.exit(not self.print_help())


def mkdir(dirname, overwrite=False, logger=None):
    if op.isdir(dirname):
        if not overwrite:
            return False
        shutil.rmtree(dirname)
    os.makedirs(dirname, exist_ok=True)
    if logger:
        logger.debug("Created folder `{}`.".format(dirname))
    return True


def listify(a):
    return a if isinstance(a, (list, tuple)) else [a]
