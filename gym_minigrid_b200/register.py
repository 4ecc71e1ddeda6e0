"""The drop-in boundary: same env ids, same `register()` contract, same `env_list` as the
reference's gym_minigrid/register.py:1-21 -- but `make()` returns a *batched* env living on a B200.

Entry points keep the reference's 'gym_minigrid.envs:<Class>' strings; each maps to the static
config the reference bakes into that class's constructor.
"""
from . import _lib

GEN_EMPTY, GEN_DOORKEY, GEN_FOURROOMS, GEN_DYNOBS, GEN_KEYCORRIDOR = range(5)

env_list = []
_specs = {}

MISSION_GOAL = "get to the green goal square"                      # empty.py:57, dynamicobstacles.py:58
MISSION_DOORKEY = "use the key to open the door and then get to the goal"   # doorkey.py:44
MISSION_FOURROOMS = "Reach the goal"                               # fourrooms.py:69
MISSION_KEYCORRIDOR = "pick up the %s %s"                          # keycorridor.py:49


def _empty(size, random_start=False):
    # EmptyEnv.__init__ (empty.py:10-28): max_steps=4*size*size, see_through_walls=True
    return dict(gen=GEN_EMPTY, width=size, height=size, max_steps=4 * size * size, see_through=1, n_actions=7,
                n_obstacles=0, room_size=0, num_rows=0, random_start=int(random_start), mission=MISSION_GOAL)


def _doorkey(size):
    # DoorKeyEnv.__init__ (doorkey.py:9-13): max_steps=10*size*size
    return dict(gen=GEN_DOORKEY, width=size, height=size, max_steps=10 * size * size, see_through=0, n_actions=7,
                n_obstacles=0, room_size=0, num_rows=0, random_start=0, mission=MISSION_DOORKEY)


def _dynobs(size, n_obstacles, random_start=False):
    # DynamicObstaclesEnv.__init__ (dynamicobstacles.py:10-33)
    n = int(n_obstacles) if n_obstacles <= size / 2 + 1 else int(size / 2)
    return dict(gen=GEN_DYNOBS, width=size, height=size, max_steps=4 * size * size, see_through=1, n_actions=3,
                n_obstacles=n, room_size=0, num_rows=0, random_start=int(random_start), mission=MISSION_GOAL,
                reward_range=(-1, 1))


def _keycorridor(room_size, num_rows):
    # KeyCorridor.__init__ (keycorridor.py:10-24) -> RoomGrid.__init__ (roomgrid.py:69-97)
    return dict(gen=GEN_KEYCORRIDOR, width=(room_size - 1) * 3 + 1, height=(room_size - 1) * num_rows + 1,
                max_steps=30 * room_size ** 2, see_through=0, n_actions=7, n_obstacles=0, room_size=room_size,
                num_rows=num_rows, random_start=0, mission=MISSION_KEYCORRIDOR)


# reference class name -> config  (envs/empty.py, doorkey.py, fourrooms.py, dynamicobstacles.py, keycorridor.py)
_CLASS_CONFIGS = {
    "EmptyEnv5x5": _empty(5), "EmptyRandomEnv5x5": _empty(5, True), "EmptyEnv6x6": _empty(6),
    "EmptyRandomEnv6x6": _empty(6, True), "EmptyEnv": _empty(8), "EmptyRandomEnv8x8": _empty(8, True),
    "EmptyEnv16x16": _empty(16),
    "DoorKeyEnv5x5": _doorkey(5), "DoorKeyEnv6x6": _doorkey(6), "DoorKeyEnv": _doorkey(8), "DoorKeyEnv16x16": _doorkey(16),
    "FourRoomsEnv": dict(gen=GEN_FOURROOMS, width=19, height=19, max_steps=500, see_through=0, n_actions=7, n_obstacles=0,
                         room_size=0, num_rows=0, random_start=0, mission=MISSION_FOURROOMS),   # fourrooms.py:14-17
    "DynamicObstaclesEnv5x5": _dynobs(5, 2), "DynamicObstaclesRandomEnv5x5": _dynobs(5, 2, True),
    "DynamicObstaclesEnv6x6": _dynobs(6, 3), "DynamicObstaclesRandomEnv6x6": _dynobs(6, 3, True),
    "DynamicObstaclesEnv": _dynobs(8, 4), "DynamicObstaclesEnv16x16": _dynobs(16, 8),
    "KeyCorridorS3R1": _keycorridor(3, 1), "KeyCorridorS3R2": _keycorridor(3, 2), "KeyCorridorS3R3": _keycorridor(3, 3),
    "KeyCorridorS4R3": _keycorridor(4, 3), "KeyCorridorS5R3": _keycorridor(5, 3), "KeyCorridorS6R3": _keycorridor(6, 3),
}
for _name, _cfg in _CLASS_CONFIGS.items():
    # minigrid.py:1263: `'v1' in self.__class__.__name__` -- true for every "...Env16x16" class
    _cfg["lava_v1"] = int("v1" in _name)
    _cfg.setdefault("reward_range", (0, 1))


def register(id, entry_point, reward_threshold=0.95):
    """Same contract as the reference (register.py:5-21): id must start with 'MiniGrid-' and be
    unique; entry_point is the reference's 'gym_minigrid.envs:<Class>' string."""
    assert id.startswith("MiniGrid-")
    assert id not in env_list
    cls = entry_point.split(":")[-1]
    if cls not in _CLASS_CONFIGS:
        raise KeyError("no B200 config for entry point %r (out of the hot-path scope, see DESIGN.md)" % entry_point)
    _specs[id] = dict(id=id, entry_point=entry_point, reward_threshold=reward_threshold, config=dict(_CLASS_CONFIGS[cls]))
    env_list.append(id)


def spec(id):
    if id not in _specs:
        raise KeyError("unknown or unsupported env id %r; supported: %s" % (id, ", ".join(env_list)))
    return _specs[id]


def make(id, num_envs=1, device=None, seed=1337, env_id_base=0, autoreset=True):
    """gym.make(id) for a batch: returns a VecMiniGridEnv with `num_envs` independent envs on one GPU."""
    from .vec_env import VecMiniGridEnv
    return VecMiniGridEnv(spec(id), num_envs=num_envs, device=device, seed=seed, env_id_base=env_id_base, autoreset=autoreset)


# ids exactly as registered by the reference env files
for _id, _cls in [
    ("MiniGrid-Empty-5x5-v0", "EmptyEnv5x5"), ("MiniGrid-Empty-Random-5x5-v0", "EmptyRandomEnv5x5"),
    ("MiniGrid-Empty-6x6-v0", "EmptyEnv6x6"), ("MiniGrid-Empty-Random-6x6-v0", "EmptyRandomEnv6x6"),
    ("MiniGrid-Empty-Random-8x8-v0", "EmptyRandomEnv8x8"), ("MiniGrid-Empty-8x8-v0", "EmptyEnv"),
    ("MiniGrid-Empty-16x16-v0", "EmptyEnv16x16"),
    ("MiniGrid-DoorKey-5x5-v0", "DoorKeyEnv5x5"), ("MiniGrid-DoorKey-6x6-v0", "DoorKeyEnv6x6"),
    ("MiniGrid-DoorKey-8x8-v0", "DoorKeyEnv"), ("MiniGrid-DoorKey-16x16-v0", "DoorKeyEnv16x16"),
    ("MiniGrid-FourRooms-v0", "FourRoomsEnv"),
    ("MiniGrid-Dynamic-Obstacles-5x5-v0", "DynamicObstaclesEnv5x5"),
    ("MiniGrid-Dynamic-Obstacles-Random-5x5-v0", "DynamicObstaclesRandomEnv5x5"),
    ("MiniGrid-Dynamic-Obstacles-6x6-v0", "DynamicObstaclesEnv6x6"),
    ("MiniGrid-Dynamic-Obstacles-Random-6x6-v0", "DynamicObstaclesRandomEnv6x6"),
    ("MiniGrid-Dynamic-Obstacles-8x8-v0", "DynamicObstaclesEnv"),
    ("MiniGrid-Dynamic-Obstacles-16x16-v0", "DynamicObstaclesEnv16x16"),
    ("MiniGrid-KeyCorridorS3R1-v0", "KeyCorridorS3R1"), ("MiniGrid-KeyCorridorS3R2-v0", "KeyCorridorS3R2"),
    ("MiniGrid-KeyCorridorS3R3-v0", "KeyCorridorS3R3"), ("MiniGrid-KeyCorridorS4R3-v0", "KeyCorridorS4R3"),
    ("MiniGrid-KeyCorridorS5R3-v0", "KeyCorridorS5R3"), ("MiniGrid-KeyCorridorS6R3-v0", "KeyCorridorS6R3"),
]:
    register(id=_id, entry_point="gym_minigrid.envs:" + _cls)
