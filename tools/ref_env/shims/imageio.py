"""imageio stand-in (test shim): offline_testing_simple.py imports it at module level (:13) and never calls it."""


def imwrite(*a, **kw):
    raise NotImplementedError("imageio stand-in: rendering is out of scope")


imread = mimsave = get_writer = imwrite
