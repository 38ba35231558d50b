"""The local `paralleltask` stand-in (compat/paralleltask): the API surface the reference's driver uses
(source/nextPolish:11,237-249,396-518) with job_type = local."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))


def test_standin_runs_jobs_marks_done_and_reruns_only_failures(tmp_path, monkeypatch):
    monkeypatch.syspath_prepend(os.path.join(ROOT, "compat"))
    sys.modules.pop("paralleltask", None)
    from paralleltask import Task
    script = tmp_path / "03.map.ref.sh"
    flag = tmp_path / "allow"
    script.write_text("echo one > out.txt\n"
                      "test -e %s && echo two > out.txt\n" % flag +       # fails until `flag` exists
                      "echo three > out.txt; echo more >> out.txt\n")
    task = Task(str(script), dir_prefix="map_genome", job_prefix="nextPolish", convert_path=False)
    assert len(task.jobs) == 3 and not task.is_finished()
    assert all(os.path.basename(j.path) == "nextPolish.sh" for j in task.jobs)
    assert len({os.path.dirname(j.path) for j in task.jobs}) == 3          # every job has its own working directory
    task.set_run(max_parallel_job=2, job_type="local", mem="3G", use_drmaa=False, submit=None, kill=None)
    total = len(task.run.unfinished_jobs)
    task.run.start()
    assert not task.run.is_finished() and len(task.run.unfinished_jobs) == 1 < total
    assert isinstance(task.run.unfinished_jobs[0].err, str)
    assert open(os.path.join(os.path.dirname(task.jobs[0].path), "out.txt")).read() == "one\n"    # ran inside its directory
    flag.write_text("")
    before = os.path.getmtime(task.jobs[0].path + ".done")
    task.run.rerun()
    assert task.run.is_finished()
    assert os.path.getmtime(task.jobs[0].path + ".done") == before         # finished jobs are not repeated
    task.set_task_finished()
    assert Task(str(script), dir_prefix="map_genome", job_prefix="nextPolish", convert_path=False).is_finished()


def test_standin_groups_lines_and_rejects_cluster_job_types(tmp_path, monkeypatch):
    monkeypatch.syspath_prepend(os.path.join(ROOT, "compat"))
    sys.modules.pop("paralleltask", None)
    from paralleltask import Task
    script = tmp_path / "x.sh"
    script.write_text("echo a > a\n\n# comment\necho b > b\necho c > c\n")
    task = Task(str(script), group=2)
    assert [len(j.lines) for j in task.jobs] == [2, 1]
    task.set_run(job_type="local")
    task.run.start()
    assert task.run.is_finished()
    try:
        task.set_run(job_type="sge")
        assert False
    except NotImplementedError:
        pass
