import gzip


def zopen(filename, mode="rt", *args, **kwargs):
    if str(filename).endswith(".gz"):
        return gzip.open(filename, mode, *args, **kwargs)
    return open(filename, mode, *args, **kwargs)
