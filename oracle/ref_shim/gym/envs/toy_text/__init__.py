from gym.envs.toy_text import frozen_lake
