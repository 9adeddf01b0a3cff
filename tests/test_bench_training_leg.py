"""bench.py's training leg (extra.training) runs tools/bench_train.py in one CHILD process per rank with its own rendezvous, so that a failure in the
collective path cannot take the headline line down.  The mechanics are tested here without a GPU: two workers launched by the real torchrun (as the
driver launches bench.py for N > 1) each call bench.time_training_step with a stand-in child that joins a gloo group of the children and all-reduces;
a failing and a hanging child must come back as 'unavailable' within the timeout."""
import json
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = textwrap.dedent('''
    import json, os, sys, time
    import torch, torch.distributed as dist
    mode = sys.argv[1]
    world, rank = int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('RANK', 0))
    if mode == 'fail' and rank == world - 1:
        sys.exit(3)
    if mode == 'hang':
        time.sleep(60)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t)
    if rank == 0:
        print(json.dumps(dict(value=float(t.item()), unit='img/s', ms_per_step=1.0, n_gpus=world, port=os.environ['MASTER_PORT'])))
    dist.destroy_process_group()
''')

PARENT = textwrap.dedent('''
    import json, os, sys
    sys.path.insert(0, {root!r})
    import torch.distributed as dist
    import bench
    world, rank = int(os.environ['WORLD_SIZE']), int(os.environ['RANK'])
    dist.init_process_group('gloo', rank=rank, world_size=world)          # the parents' own group stays up while the children run, as in bench.py
    out = {{}}
    for mode, timeout in (('ok', 60), ('fail', 8), ('hang', 5)):
        r = bench.time_training_step(world, rank, timeout=timeout, cmd=[sys.executable, {child!r}, mode])
        dist.barrier()
        out[mode] = r
    if rank == 0:
        print('RESULT ' + json.dumps(dict(out=out, parent_port=os.environ['MASTER_PORT'])))
    else:
        assert all(v is None for v in out.values())
    dist.destroy_process_group()
''')


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def test_training_leg_children_rendezvous_under_torchrun(tmp_path):
    child, parent = tmp_path / 'child.py', tmp_path / 'parent.py'
    child.write_text(CHILD)
    parent.write_text(PARENT.format(root=ROOT, child=str(child)))
    for attempt in range(3):                                 # (the children use port + 17: on the rare collision with a foreign listener, take another port)
        port = _free_port()
        r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                            '--master-port', str(port), str(parent)], capture_output=True, text=True, timeout=300, cwd=ROOT)
        assert r.returncode == 0, r.stderr[-2000:]
        line = [ln for ln in r.stdout.splitlines() if ln.startswith('RESULT ')][-1]
        res = json.loads(line[len('RESULT '):])
        if 'unavailable' not in res['out']['ok']:
            break
    ok = res['out']['ok']
    assert ok['value'] == 3.0 and ok['n_gpus'] == 2                               # 1 + 2 summed across the two children
    assert res['out']['fail'] is not None and 'unavailable' in res['out']['fail']
    assert 'timed out' in res['out']['hang']['unavailable']


def test_training_leg_single_process(tmp_path):
    sys.path.insert(0, ROOT)
    import bench
    child = tmp_path / 'child.py'
    child.write_text(CHILD)
    env = {k: os.environ.pop(k) for k in ('WORLD_SIZE', 'RANK', 'MASTER_PORT') if k in os.environ}
    try:
        os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(_free_port())
        r = bench.time_training_step(1, 0, timeout=60, cmd=[sys.executable, str(child), 'ok'])
    finally:
        os.environ.update(env)
    assert r['value'] == 1.0 and r['n_gpus'] == 1
