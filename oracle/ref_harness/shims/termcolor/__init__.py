"""Import shim: the reference's logger imports termcolor, which this image lacks."""


def colored(text, *args, **kwargs):
    return text
