"""Local stand-in for the `paralleltask` package (moold/ParallelTask) that the reference's workflow driver imports
(`from paralleltask import Task`, source/nextPolish:11).  The reference does not pin a version and the package is not
vendored; no arithmetic lives in it.  Only what the driver uses is provided, for `job_type = local`
(source/nextPolish:237-249,396-518):

    task = Task(path_of_a_shell_script, dir_prefix=..., job_prefix=..., convert_path=..., group=...)
    task.jobs[i].path            # the job's own script; its directory is the job's working directory
    task.is_finished(); task.set_task_finished()
    task.set_run(max_parallel_job=..., job_type='local', ...)   # every other keyword is accepted and ignored
    task.run.start(); task.run.is_finished(); task.run.unfinished_jobs[i].err; task.run.rerun()

A task script holds one command per line (`group` lines per job).  Job i runs in `<script>.work/<dir_prefix><i>/`, from a
generated `<job_prefix>.sh` that does `set -e; cd <job dir>; <command lines>`; stdout / stderr go to `<job_prefix>.sh.o` /
`.e`, success leaves `<job_prefix>.sh.done`, and a finished task leaves `<script>.done` — so a re-run of the driver skips
finished tasks and `rerun()` only repeats the jobs without a done marker.

Use: put this directory's parent on PYTHONPATH (`PYTHONPATH=<repo>/compat python <NextPolish>/nextPolish run.cfg`)."""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

__all__ = ["Task"]


class Job:
    def __init__(self, path, lines):
        self.path = path                      # the job script
        self.lines = lines
        self.out, self.err, self.done = path + ".o", path + ".e", path + ".done"
        self.returncode = None

    def write(self):
        d = os.path.dirname(self.path)
        os.makedirs(d, exist_ok=True)
        with open(self.path, "w") as f:
            f.write("#!/bin/sh\nset -e\ncd %s\n" % _quote(d))
            for ln in self.lines:
                f.write(ln.rstrip("\n") + "\n")

    def is_finished(self):
        return os.path.exists(self.done)

    def execute(self):
        with open(self.out, "w") as o, open(self.err, "w") as e:
            self.returncode = subprocess.call(["sh", self.path], stdout=o, stderr=e)
        if self.returncode == 0:
            open(self.done, "w").close()
        return self.returncode


def _quote(s):
    return "'" + s.replace("'", "'\\''") + "'"


class Run:
    def __init__(self, jobs, max_parallel_job=1):
        self.jobs = jobs
        self.max_parallel_job = max(1, int(max_parallel_job or 1))

    @property
    def unfinished_jobs(self):
        return [j for j in self.jobs if not j.is_finished()]

    def is_finished(self):
        return not self.unfinished_jobs

    def start(self):
        todo = self.unfinished_jobs
        if not todo:
            return
        with ThreadPoolExecutor(max_workers=min(self.max_parallel_job, len(todo))) as pool:
            list(pool.map(lambda j: j.execute(), todo))

    def rerun(self):
        self.start()


class Task:
    def __init__(self, path, dir_prefix="task", job_prefix="job", convert_path=True, group=1, **_ignored):
        self.path = os.path.abspath(path) if convert_path else path
        self.done = self.path + ".done"
        with open(self.path) as f:
            lines = [ln for ln in f.read().split("\n") if ln.strip() and not ln.lstrip().startswith("#")]
        group = max(1, int(group or 1))
        work = self.path + ".work"
        self.jobs = []
        for i in range(0, len(lines), group):
            d = os.path.join(work, "%s%d" % (dir_prefix, i // group))
            job = Job(os.path.join(d, job_prefix + ".sh"), lines[i:i + group])
            job.write()
            self.jobs.append(job)
        self.run = None

    def is_finished(self):
        return os.path.exists(self.done)

    def set_task_finished(self):
        open(self.done, "w").close()

    def set_run(self, max_parallel_job=1, job_type="local", **_ignored):
        if job_type not in ("local", None):
            raise NotImplementedError("paralleltask stand-in: only job_type = local (got %r)" % (job_type,))
        self.run = Run(self.jobs, max_parallel_job)
        return self.run
