"""Kinematic stand-in for the PyBullet calls the TIP reference makes -- TEST INFRASTRUCTURE (row N2).

The reference's `SimAgent` (bullet_agent.py:18-157) uses PyBullet purely as a forward-kinematics engine when
`kinematic_only=True` (render_funcs.py:100-108): it loads `data/amass.urdf`, sets the base pose and the
spherical joint quaternions (`resetBasePositionAndOrientation`, `resetJointStatesMultiDof`) and reads link
frames back (`getLinkStates`, bullet_utils.py:79-157).  This module restates exactly those semantics in numpy:

* links are numbered in URDF order after the base (URDF_MAINTAIN_LINK_ORDER); joint i moves link i;
* world link frame  W_i = W_parent * T(joint origin xyz/rpy) * R(q_i)   (q_i = identity for fixed joints);
* `getLinkStates` item [4],[5] = URDF link frame, [0],[1] = centre-of-mass frame = W_i * T(inertial origin),
  [2],[3] = local inertial frame; quaternions are xyzw; the base position is the base's COM frame;
* dynamics / collision / rendering calls are accepted and ignored.

It is not a physics engine and is never imported by the product.
"""
import xml.etree.ElementTree as ET

import numpy as np
from scipy.spatial.transform import Rotation


class error(Exception):
    pass


# connection modes / flags / enums (values as in pybullet)
SHARED_MEMORY, DIRECT, GUI = 1, 2, 7
JOINT_REVOLUTE, JOINT_PRISMATIC, JOINT_SPHERICAL, JOINT_PLANAR, JOINT_FIXED = 0, 1, 2, 3, 4
URDF_USE_SELF_COLLISION = 8
URDF_USE_SELF_COLLISION_EXCLUDE_PARENT = 16
URDF_USE_SELF_COLLISION_EXCLUDE_ALL_PARENTS = 32
URDF_MAINTAIN_LINK_ORDER = 512
ACTIVATION_STATE_ENABLE_SLEEPING, ACTIVATION_STATE_DISABLE_SLEEPING = 1, 2
ACTIVATION_STATE_WAKE_UP, ACTIVATION_STATE_SLEEP = 4, 8
ACTIVATION_STATE_ENABLE_WAKEUP, ACTIVATION_STATE_DISABLE_WAKEUP = 16, 32
VELOCITY_CONTROL, TORQUE_CONTROL, POSITION_CONTROL = 0, 1, 2
COV_ENABLE_GUI, COV_ENABLE_SHADOWS, COV_ENABLE_WIREFRAME = 1, 2, 3
COV_ENABLE_RENDERING, COV_ENABLE_RGB_BUFFER_PREVIEW = 7, 13
COV_ENABLE_DEPTH_BUFFER_PREVIEW, COV_ENABLE_SEGMENTATION_MARK_PREVIEW = 14, 15
GEOM_SPHERE, GEOM_BOX, GEOM_HEIGHTFIELD = 2, 3, 9

_bodies = {}
_next_body = [0]


def _rpy_to_R(rpy):
    # URDF fixed-axis roll/pitch/yaw = R_z(yaw) R_y(pitch) R_x(roll)
    return Rotation.from_euler("xyz", rpy).as_matrix()


def _origin(elem):
    xyz, rpy = np.zeros(3), np.zeros(3)
    if elem is not None:
        o = elem.find("origin")
        if o is not None:
            xyz = np.array([float(v) for v in o.get("xyz", "0 0 0").split()])
            rpy = np.array([float(v) for v in o.get("rpy", "0 0 0").split()])
    T = np.eye(4)
    T[:3, :3] = _rpy_to_R(rpy)
    T[:3, 3] = xyz
    return T


class _Body:
    def __init__(self, path, scale):
        root = ET.parse(path).getroot()
        links = root.findall("link")
        joints = root.findall("joint")
        names = [l.get("name") for l in links]
        child_of = {j.find("child").get("link"): j for j in joints}
        base = [n for n in names if n not in child_of]
        assert len(base) == 1, "expected one base link"
        self.base_name = base[0]
        order = [n for n in names if n != self.base_name]          # URDF_MAINTAIN_LINK_ORDER
        self.link_names = order
        self.index = {n: i for i, n in enumerate(order)}
        self.index[self.base_name] = -1
        self.n = len(order)
        link_by_name = {l.get("name"): l for l in links}
        self.inertial = {}                                          # link index -> T(link frame -> COM frame)
        self.mass = {}
        for n in names:
            ine = link_by_name[n].find("inertial")
            T = _origin(ine)
            T[:3, 3] *= scale
            self.inertial[self.index[n]] = T
            m = ine.find("mass") if ine is not None else None
            self.mass[self.index[n]] = float(m.get("value")) if m is not None else 0.0
        self.parent, self.origin, self.jtype, self.jname, self.axis = [], [], [], [], []
        types = {"revolute": JOINT_REVOLUTE, "continuous": JOINT_REVOLUTE, "prismatic": JOINT_PRISMATIC,
                 "spherical": JOINT_SPHERICAL, "fixed": JOINT_FIXED, "planar": JOINT_PLANAR}
        for n in order:
            j = child_of[n]
            self.parent.append(self.index[j.find("parent").get("link")])
            T = _origin(j)
            T[:3, 3] *= scale
            self.origin.append(T)
            self.jtype.append(types[j.get("type")])
            self.jname.append(j.get("name"))
            ax = j.find("axis")
            self.axis.append(np.array([float(v) for v in ax.get("xyz").split()]) if ax is not None
                             else np.zeros(3))
        # state
        self.base_p = np.zeros(3)
        self.base_Q = np.array([0.0, 0, 0, 1])
        self.base_v = np.zeros(3)
        self.base_w = np.zeros(3)
        self.q = [np.array([0.0, 0, 0, 1]) if t == JOINT_SPHERICAL else (np.zeros(1) if t == JOINT_REVOLUTE else np.zeros(0))
                  for t in self.jtype]
        self.dq = [np.zeros(3) if t == JOINT_SPHERICAL else (np.zeros(1) if t == JOINT_REVOLUTE else np.zeros(0))
                   for t in self.jtype]

    def joint_R(self, i):
        t = self.jtype[i]
        if t == JOINT_SPHERICAL:
            return Rotation.from_quat(self.q[i]).as_matrix()
        if t == JOINT_REVOLUTE:
            return Rotation.from_rotvec(self.axis[i] * float(self.q[i][0])).as_matrix()
        return np.eye(3)

    def fk(self):
        """world transforms of every URDF link frame (index -1 = base)."""
        Wcom = np.eye(4)
        Wcom[:3, :3] = Rotation.from_quat(self.base_Q).as_matrix()
        Wcom[:3, 3] = self.base_p
        W = {-1: Wcom @ np.linalg.inv(self.inertial[-1])}
        for i in range(self.n):                                     # parents precede children in URDF order
            if self.parent[i] not in W:
                raise error("URDF link order: parent after child")
            J = np.eye(4)
            J[:3, :3] = self.joint_R(i)
            W[i] = W[self.parent[i]] @ self.origin[i] @ J
        return W


def _body(bid):
    if bid not in _bodies:
        raise error("unknown body id %r" % (bid,))
    return _bodies[bid]


def _pQ(T):
    return tuple(T[:3, 3]), tuple(Rotation.from_matrix(T[:3, :3]).as_quat())


# ---- connection ----------------------------------------------------------------------------------
def connect(mode, options="", **kw):
    return -1 if mode == SHARED_MEMORY else 0


def disconnect(physicsClientId=0):
    pass


def resetSimulation(physicsClientId=0):
    _bodies.clear()


def isNumpyEnabled():
    return False


# ---- bodies --------------------------------------------------------------------------------------
def loadURDF(fileName, basePosition=(0, 0, 0), baseOrientation=(0, 0, 0, 1), globalScaling=1.0,
             useFixedBase=False, flags=0, physicsClientId=0, **kw):
    b = _Body(fileName, float(globalScaling))
    b.base_p = np.asarray(basePosition, dtype=float)
    b.base_Q = np.asarray(baseOrientation, dtype=float)
    bid = _next_body[0]
    _next_body[0] += 1
    _bodies[bid] = b
    return bid


def getNumJoints(bodyUniqueId, physicsClientId=0):
    return _body(bodyUniqueId).n


def getJointInfo(bodyUniqueId, jointIndex, physicsClientId=0):
    b = _body(bodyUniqueId)
    i = jointIndex
    # joint frame relative to the parent's inertial (COM) frame
    Tp = np.linalg.inv(b.inertial[b.parent[i]]) @ b.origin[i]
    p, Q = _pQ(Tp)
    qi = {JOINT_SPHERICAL: 7, JOINT_REVOLUTE: 7}.get(b.jtype[i], -1)
    return (i, b.jname[i].encode(), b.jtype[i], qi, qi, 0, 0.0, 0.0, 0.0, -1.0, 0.0, 0.0,
            b.link_names[i].encode(), tuple(b.axis[i]), p, Q, b.parent[i])


def getDynamicsInfo(bodyUniqueId, linkIndex, physicsClientId=0):
    b = _body(bodyUniqueId)
    p, Q = _pQ(b.inertial[linkIndex])
    return (b.mass[linkIndex], 0.5, (0.0, 0.0, 0.0), p, Q, 0.0, 0.0, 0.0, -1.0, -1.0, 2, 0.001)


def resetBasePositionAndOrientation(bodyUniqueId, posObj, ornObj, physicsClientId=0):
    b = _body(bodyUniqueId)
    b.base_p = np.asarray(posObj, dtype=float).copy()
    b.base_Q = np.asarray(ornObj, dtype=float).copy()


def resetBaseVelocity(objectUniqueId, linearVelocity=None, angularVelocity=None, physicsClientId=0):
    b = _body(objectUniqueId)
    if linearVelocity is not None:
        b.base_v = np.asarray(linearVelocity, dtype=float).copy()
    if angularVelocity is not None:
        b.base_w = np.asarray(angularVelocity, dtype=float).copy()


def getBasePositionAndOrientation(bodyUniqueId, physicsClientId=0):
    b = _body(bodyUniqueId)
    return tuple(b.base_p), tuple(b.base_Q)


def getBaseVelocity(bodyUniqueId, physicsClientId=0):
    b = _body(bodyUniqueId)
    return tuple(b.base_v), tuple(b.base_w)


def resetJointStatesMultiDof(bodyUniqueId, jointIndices, targetValues, targetVelocities=None, physicsClientId=0):
    b = _body(bodyUniqueId)
    for k, j in enumerate(jointIndices):
        v = np.asarray(targetValues[k], dtype=float)
        if b.jtype[j] == JOINT_SPHERICAL:
            assert v.shape == (4,), "spherical joints take xyzw quaternions"
            b.q[j] = v / np.linalg.norm(v)
        elif b.jtype[j] == JOINT_REVOLUTE:
            b.q[j] = v.reshape(1)
        if targetVelocities is not None and len(b.dq[j]):
            b.dq[j] = np.asarray(targetVelocities[k], dtype=float).reshape(b.dq[j].shape)


def resetJointStateMultiDof(bodyUniqueId, jointIndex, targetValue, targetVelocity=None, physicsClientId=0):
    resetJointStatesMultiDof(bodyUniqueId, [jointIndex], [targetValue],
                             None if targetVelocity is None else [targetVelocity])


def getJointStatesMultiDof(bodyUniqueId, jointIndices, physicsClientId=0):
    b = _body(bodyUniqueId)
    return tuple((tuple(b.q[j]), tuple(b.dq[j]), (0.0,) * 6, (0.0,) * len(b.dq[j])) for j in jointIndices)


def getJointStateMultiDof(bodyUniqueId, jointIndex, physicsClientId=0):
    return getJointStatesMultiDof(bodyUniqueId, [jointIndex])[0]


def getLinkStates(bodyUniqueId, linkIndices, computeLinkVelocity=0, computeForwardKinematics=0, physicsClientId=0):
    b = _body(bodyUniqueId)
    W = b.fk()
    out = []
    for i in linkIndices:
        com = W[i] @ b.inertial[i]
        p_com, Q_com = _pQ(com)
        p_in, Q_in = _pQ(b.inertial[i])
        p_f, Q_f = _pQ(W[i])
        out.append((p_com, Q_com, p_in, Q_in, p_f, Q_f, (0.0, 0.0, 0.0), (0.0, 0.0, 0.0)))
    return tuple(out)


def getLinkState(bodyUniqueId, linkIndex, computeLinkVelocity=0, computeForwardKinematics=0, physicsClientId=0):
    return getLinkStates(bodyUniqueId, [linkIndex])[0]


# ---- accepted and ignored (dynamics, collision, rendering) ---------------------------------------
def _noop(*a, **kw):
    return None


setCollisionFilterPair = setCollisionFilterGroupMask = changeDynamics = changeVisualShape = _noop
setJointMotorControl2 = setJointMotorControlMultiDof = enableJointForceTorqueSensor = _noop
configureDebugVisualizer = resetDebugVisualizerCamera = removeAllUserDebugItems = _noop
setGravity = setTimeStep = stepSimulation = setRealTimeSimulation = setPhysicsEngineParameter = _noop
removeBody = removeUserDebugItem = _noop


def createVisualShape(*a, **kw):
    return -1


def createCollisionShape(*a, **kw):
    return -1


class _Marker:
    """A free body without links (the SBP marker spheres of render_funcs.py:213-225): only its base pose is kept."""
    n = 0

    def __init__(self, pos, orn):
        self.base_p = np.asarray(pos, dtype=float)
        self.base_Q = np.asarray(orn, dtype=float)
        self.base_v = np.zeros(3)
        self.base_w = np.zeros(3)


def createMultiBody(baseMass=0.0, baseCollisionShapeIndex=-1, baseVisualShapeIndex=-1, basePosition=(0, 0, 0),
                    baseOrientation=(0, 0, 0, 1), *a, **kw):
    bid = _next_body[0]
    _next_body[0] += 1
    _bodies[bid] = _Marker(basePosition, baseOrientation)
    return bid


def loadTexture(*a, **kw):
    return -1


def addUserDebugLine(*a, **kw):
    return -1


def addUserDebugText(*a, **kw):
    return -1


def getVisualShapeData(*a, **kw):
    return ()
