"""Reference-arm support: the recipe that ships the unmodified reference to the GPU box (install_reference.py) and the
import shim that lets it run on this image's library versions (refshim.py).  Test / bench infrastructure only:
nothing under stemseg_b200/ imports this package."""
