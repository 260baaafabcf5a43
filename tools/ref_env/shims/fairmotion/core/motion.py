"""fairmotion.core.motion placeholder (test shim): the TIP runners only need the names to import."""


class Pose(object):
    def __init__(self, skel=None, data=None):
        self.skel, self.data = skel, data


class Motion(object):
    def __init__(self, name="motion", skel=None, fps=60):
        self.name, self.skel, self.fps, self.poses = name, skel, fps, []
