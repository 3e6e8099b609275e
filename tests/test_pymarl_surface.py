"""GPU suite: the PyMARL MultiAgentEnv facade behaves like the reference surface and can be
driven by the reference runner's exact call pattern (src/runners/episode_runner.py:48-141), including the
th.tensor(list-of-numpy) conversion of EpisodeBatch.update (src/components/episode_buffer.py:102-109)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _env(name="cleanup", **kw):
    from homophily_marl_b200 import REGISTRY
    args = dict(num_agents=3, render=False, episode_limit=12, is_replay=False, view_size=7, map="default3", seed=5,
                extra_args=dict(random_spawn_point=False, random_spawn_rotation=0, disable_rotation_action=True,
                                disable_fire_action=True, obs_color="simplified"), quiet=True)
    args.update(kw)
    return REGISTRY[name](**args)


def test_env_info_before_reset_and_static_queries():
    env = _env()
    info = env.get_env_info()
    assert info == {"state_shape": (3, 10, 10), "obs_shape": (3, 15, 15), "n_actions": 9, "n_agents": 3,
                    "episode_limit": 12, "units_type_id": None, "own_feature_size": None,
                    "state_dims": (10, 10), "obs_dims": (15, 15)}
    assert env.get_avail_actions() == [[1, 1, 1, 1, 1, 0, 0, 0, 1]] * 3
    assert env.get_total_actions() == 9 and env.episode_limit == 12 and env.n_agents == 3
    assert env.reset() is None and env.seed() is None and env.get_stats() == {}
    h = _env("harvest", num_agents=5, map="default10", view_size=15)
    assert h.get_env_info()["n_actions"] == 8 and h.get_env_info()["obs_shape"] == (3, 31, 31)
    assert h.get_avail_agent_actions(0) == [1, 1, 1, 1, 1, 0, 0, 0]


def test_runner_call_pattern_and_types_match_oracle():
    from oracle import oracle as O
    env = _env()
    ora = O.OracleBatch.from_spec(env.sim.spec, n_envs=1, seed=5)
    env.reset()
    ora.reset()
    rs = np.random.RandomState(0)
    terminated, t, episode_return = False, 0, 0
    dev = env.sim.device
    while not terminated:
        pre = {"state": [env.get_state()], "avail_actions": [env.get_avail_actions()], "obs": [env.get_obs()],
               "agent_pos": [env.get_agent_pos()], "agent_orientation": [env.get_agent_orientation()]}
        tens = {k: torch.tensor(v, dtype=torch.float32, device=dev) for k, v in pre.items()}      # EpisodeBatch.update
        assert tens["obs"].shape == (1, 3, 3, 15, 15) and tens["state"].shape == (1, 3, 10, 10)
        assert np.array_equal(np.stack(pre["obs"][0]) * 256, ora.obs_one(0))
        assert np.array_equal(pre["state"][0] * 256, ora.state_one(0))
        assert np.array_equal(pre["agent_pos"][0], ora.pos_rc[0].astype(float))
        actions = torch.as_tensor(rs.randint(0, 9, size=(1, 3, 1)), device=dev)                    # LongTensor [bs, n, 1]
        reward, terminated, info = env.step(actions[0])
        r2, c2, cnt, done = ora.step_one(actions[0, :, 0].cpu().numpy())
        assert isinstance(terminated, bool) and terminated == done
        assert reward.dtype == np.float64 and np.array_equal(reward, r2.astype(float))
        assert np.array_equal(info["clean_num"], c2.astype(float)) and np.allclose(info["apple_den"], cnt / 100)
        torch.tensor([(terminated,)], dtype=torch.uint8)                                            # SURVEY 8b pitfall
        torch.tensor([(reward,)], dtype=torch.float32)
        episode_return += reward
        t += 1
    assert t == 12 and set(info) == {"collective_return", "equality_metric", "clean_num", "apple_den"}
    R = ora.ep_ret[0].astype(float)
    assert info["collective_return"] == R.sum()
    want = 1.0 if R.sum() == 0 else 1 - np.abs(R.reshape(1, -1) - R.reshape(-1, 1)).sum() / (2 * len(R) * np.abs(R).sum())
    assert info["equality_metric"] == want
    last_obs = env.get_obs()
    assert len(last_obs) == 3 and last_obs[0].shape == (3, 15, 15) and last_obs[0].dtype == np.float64
    env.reset()
    assert env.rewards is None and env._episode_steps == 0
    env.close()


def test_invalid_action_raises_keyerror_like_action_map():
    env = _env("harvest", num_agents=2, map="default10", view_size=7)
    env.reset()
    with pytest.raises(KeyError):
        env.step([8, 0])                                   # CLEAN does not exist in Harvest (agent.py:176)
    with pytest.raises(KeyError):
        env.step([0, 9])
