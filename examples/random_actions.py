"""The reference's examples/random_actions.py on the batched simulator: same loop, same types (float64 observation, float
reward, bool done, info dict), one environment stepping on the GPU through the C ABI."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import assistive_vr_gym_b200 as assistive_gym      # stands where `import gym, assistive_gym` stands in the reference

env = assistive_gym.make('ScratchItchJaco-v0')
env.render()
observation = env.reset()
total = 0.0
for t in range(200):
    env.render()
    observation, reward, done, info = env.step(env.action_space.sample())
    total += reward
print('return', total, 'task_success', info['task_success'], 'observation', observation.shape, observation.dtype)
env.close()
