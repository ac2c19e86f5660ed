"""``FitPayne`` with the reference's input dictionary and output file, driving the lock-step batched
nested sampler instead of dynesty.

Mirror of ``Payne/fitting/fitstar.py``: ``run(inputdict=...)`` (:19-217) turns the user dictionary
(``spec`` / ``phot`` / ``priordict`` / ``sampler`` / ``output`` / ``specANNpath`` / ``photANNpath`` /
``NNtype`` / ``photscale``) into ``fitargs``, the ordered parameter list with its on/off switches and
the run booleans; ``run_dynesty`` (:236-258) builds the prior and likelihood objects;
``_runsampler`` (:260-463) samples and writes one line per dead point -- iteration, the parameters,
``log(lk) log(vol) log(wt) h nc log(z) delta(log(z))`` -- then the final live points.

What differs, on purpose:
  * the sampler is ``nested.BatchedNestedSampler``: every likelihood evaluation is part of one
    ``likelihood.lnlike_batch`` call of ``queue_size`` (default ``npoints``) proposals.  The sampler keys
    of the reference keep their meaning (``npoints``, ``walks``, ``delta_logz_final``, ``flushnum``,
    ``maxiter``, ``maxcall``, ``reflective``); ``samplemethod`` must be 'rwalk' (the demo's choice,
    demo/runPayne.py:112), ``samplerbounds`` is a single ellipsoid, ``samplertype`` 'Static';
  * the parameters written for a dead point are that point's own (the reference writes
    ``likeobj.parsdict``, i.e. whatever vector the likelihood saw last, fitstar.py:347);
  * ``inputdict['precision']`` / ``['device']`` select the CUDA precision mode and device.
"""
from __future__ import annotations

import sys
from datetime import datetime

import numpy as np

from .batching import BatchedLnProb
from .fitutils import airtovacuum
from .nested import BatchedNestedSampler

_ALLPARS = ['Teff', 'log(g)', '[Fe/H]', '[a/Fe]', 'Vrad', 'Vrot', 'Vmic', 'Inst_R', 'log(R)', 'Dist', 'log(A)',
            'Av', 'Rv', 'CarbonScale']


class FitPayne(object):
    def __init__(self, **kwargs):
        from .likelihood import likelihood
        from .prior import prior
        self.prior = prior
        self.likelihood = likelihood

    # ------------------------------------------------------------------ input dictionary -> fit setup
    def run(self, *args, **kwargs):
        self.verbose = kwargs.get('verbose', True)
        if 'inputdict' not in kwargs:
            print('NO USER DEFINED INPUT DICT, NOTHING TO FIT!')
            raise IOError
        inputdict = kwargs['inputdict']
        self.priordict = inputdict.get('priordict', {})
        self.output = inputdict.get('output', 'Test.dat')
        self.samplerdict = inputdict.get('sampler', {})
        self.precision = inputdict.get('precision', 'parity')
        self.device = inputdict.get('device', None)
        fa = self.fitargs = {}
        self.spec_bool = self.phot_bool = self.modpoly_bool = self.photscale_bool = self.carbon_bool = False
        self.fitpars = list(_ALLPARS)
        self.fitpars_bool = {pp: False for pp in self.fitpars}

        if 'spec' in inputdict:                                            # fitstar.py:70-154
            sp = inputdict['spec']
            self.spec_bool = True
            fa['obs_wave'], fa['obs_flux'], fa['obs_eflux'] = sp['obs_wave'], sp['obs_flux'], sp['obs_eflux']
            fa['specANNpath'] = inputdict.get('specANNpath', None)
            fa['NNtype'] = inputdict.get('NNtype', 'LinNet')
            sel = slice(None)
            if 'wave_minmax' in sp:
                fa['wave_minmax'] = sp['wave_minmax']
                sel = (np.asarray(fa['obs_wave']) >= sp['wave_minmax'][0]) & (np.asarray(fa['obs_wave']) <= sp['wave_minmax'][1])
            for k in ['wave', 'flux', 'eflux']:
                fa['obs_%s_fit' % k] = np.asarray(fa['obs_' + k], dtype=np.float64)[sel]
            if sp.get('convertair', True):
                fa['obs_wave_fit'] = airtovacuum(fa['obs_wave_fit'])
            on = ['Teff', 'log(g)', '[Fe/H]', '[a/Fe]', 'Vrad', 'Vrot', 'Inst_R']
            if fa['NNtype'] == 'YST2':
                on.append('Vmic')
            for pp in on:
                self.fitpars_bool[pp] = True
            if sp.get('modpoly', False):                                   # fitstar.py:110-147
                self.modpoly_bool = True
                if 'blaze_coeff' in self.priordict:
                    self.polycoefarr = self.priordict['blaze_coeff']
                elif 'polyorder' in sp:
                    sig = sp.get('polysigma', 1.0)
                    self.polycoefarr = [[0.0, sig] for _ in range(sp['polyorder'] + 1)]
                else:
                    self.polycoefarr = [[0.0, 1.0] for _ in range(3)]
                self.priordict['blaze_coeff'] = self.polycoefarr
                self.polyorder = len(self.polycoefarr)
                fa['norm_polyorder'] = self.polyorder
                for ii in range(self.polyorder):
                    self.fitpars.append('pc_{}'.format(ii))
                    self.fitpars_bool['pc_{}'.format(ii)] = True

        if 'phot' in inputdict:                                            # fitstar.py:157-190
            self.phot_bool = True
            fa['photANNpath'] = inputdict.get('photANNpath', None)
            for pp in ['Teff', 'log(g)', '[Fe/H]', '[a/Fe]', 'Av']:
                self.fitpars_bool[pp] = True
            fa['obs_phot'] = {kk: inputdict['phot'][kk] for kk in inputdict['phot'].keys()}
            self.photscale_bool = inputdict.get('photscale', False)
            if self.photscale_bool:
                self.fitpars_bool['log(A)'] = True
            else:
                self.fitpars_bool['log(R)'] = True
                self.fitpars_bool['Dist'] = True
            if inputdict.get('Rvfree', False):
                raise NotImplementedError('Rvfree: the likelihood never forwards Rv (likelihood.py:103-106)')

        fa['fixedpars'] = {}                                               # fitstar.py:193-198
        for kk in self.priordict.keys():
            if isinstance(self.priordict[kk], dict) and 'fixed' in self.priordict[kk]:
                fa['fixedpars'][kk] = self.priordict[kk]['fixed']
                self.fitpars_bool[kk] = False
        return self({'fitargs': fa, 'fitpars': [self.fitpars, self.fitpars_bool], 'sampler': self.samplerdict,
                     'priordict': self.priordict,
                     'runbools': [self.spec_bool, self.phot_bool, self.modpoly_bool, self.photscale_bool,
                                  self.carbon_bool]})

    def __call__(self, indicts):
        return self.run_dynesty(indicts)

    def run_dynesty(self, indicts):
        """Name kept from the reference (fitstar.py:236-258); the sampler is the batched one."""
        fitargs, fitpars, runbools = indicts['fitargs'], indicts['fitpars'], indicts['runbools']
        self.ndim = sum(1 for pp in fitpars[0] if fitpars[1][pp])
        self.priorobj = self.prior(fitargs, indicts['priordict'], fitpars, runbools)
        self.likeobj = self.likelihood(fitargs, fitpars, runbools, verbose=False,
                                       precision=getattr(self, 'precision', 'parity'),
                                       device=getattr(self, 'device', None))
        kind = indicts['sampler'].get('samplertype', 'Static')
        if kind != 'Static':
            raise NotImplementedError("samplertype %r: only the static sampler is batched" % kind)
        return self._runsampler(indicts['sampler'])

    # ------------------------------------------------------------------ output file (fitstar.py:219-227)
    def _initoutput(self, parnames):
        self.outff = open(self.output, 'w')
        self.outff.write('Iter ')
        for pp in parnames:
            self.outff.write('{} '.format(pp))
        self.outff.write('log(lk) log(vol) log(wt) h nc log(z) delta(log(z))')
        self.outff.write('\n')

    def _writerow(self, it, vstar, parnames, tail):
        pd = {pp: vv for pp, vv in zip(self.likeobj.fitpars_i, vstar)}
        pd.update(self.likeobj.fixedpars)
        self.outff.write('{0} '.format(it))
        self.outff.write(' '.join([str(pd[q]) for q in parnames]))
        self.outff.write(' {0} {1} {2} {3} {4} {5} {6} '.format(*tail))
        self.outff.write('\n')

    def _runsampler(self, samplerdict):
        npoints = samplerdict.get('npoints', 200)
        samplemethod = samplerdict.get('samplemethod', 'rwalk')
        if samplemethod != 'rwalk':
            raise NotImplementedError("samplemethod %r: the batched sampler implements 'rwalk'" % samplemethod)
        delta_logz_final = samplerdict.get('delta_logz_final', 0.01)
        flushnum = samplerdict.get('flushnum', 10)
        numwalks = samplerdict.get('walks', 25)
        maxiter = samplerdict.get('maxiter', sys.maxsize)
        maxcall = samplerdict.get('maxcall', sys.maxsize)
        reflective = [ii for ii, par in enumerate(self.likeobj.fitpars_i) if par in samplerdict.get('reflective', [])]
        starttime = datetime.now()
        if self.verbose:
            print('Batched static nested sampler w/ rwalk, {0} walks, {1} live points, queue of {2}, Ndim = {3}, '
                  'dlog(z) = {4}: {5}'.format(numwalks, npoints, samplerdict.get('queue_size', npoints), self.ndim,
                                              delta_logz_final, starttime))
        lnprob = BatchedLnProb(self.likeobj, self.priorobj)
        sampler = BatchedNestedSampler(lnprob.batch, self.priorobj.priortrans_batch, self.ndim, nlive=npoints,
                                       walks=numwalks, queue_size=samplerdict.get('queue_size', None),
                                       reflective=reflective, seed=samplerdict.get('seed', None))
        parnames = list(self.likeobj.fitpars_i) + list(self.likeobj.fixedpars.keys())
        self._initoutput(parnames)
        ncall, nit = 0, 0
        for it, res in enumerate(sampler.sample(dlogz=delta_logz_final, maxiter=maxiter, maxcall=maxcall)):
            (worst, ustar, vstar, loglstar, logvol, logwt, logz, logzvar, h, nc, worst_it, propidx, propiter, eff,
             delta_logz) = res
            self._writerow(it, vstar, parnames, (loglstar, logvol, logwt, h, nc, logz, delta_logz))
            ncall += nc
            nit = it
            if (it % flushnum) == 0 or it == maxiter:
                self.outff.flush()
                if self.verbose and (it % (50 * flushnum)) == 0:
                    logzerr = np.sqrt(logzvar) if logzvar > 0.0 else np.nan
                    sys.stdout.write('iter: {0:d} | nc: {1:d} | ncall: {2:d} | eff(%): {3:6.3f} | logz: {4:6.3f} +/- '
                                     '{5:6.3f} | loglk: {6:6.3f} | dlogz: {7:6.3f} > {8:6.3f}\n'.format(
                                         nit, nc, sampler.ncall, eff, logz, logzerr, loglstar, delta_logz,
                                         delta_logz_final))
                    sys.stdout.flush()
            if it == maxiter:
                break
        for it2, res in enumerate(sampler.add_live_points()):
            (worst, ustar, vstar, loglstar, logvol, logwt, logz, logzvar, h, nc, worst_it, boundidx, bounditer, eff,
             delta_logz) = res
            self._writerow(nit + it2, vstar, parnames, (loglstar, logvol, logwt, h, nc, logz, delta_logz))
        self.outff.close()
        if self.verbose:
            print('RUN TIME: {0}  ({1} likelihood evaluations in {2} batched calls)'.format(
                datetime.now() - starttime, sampler.ncall, len(sampler.batch_sizes)))
        self.likeobj._set_parsdict([float(v) for v in vstar])
        return sampler
